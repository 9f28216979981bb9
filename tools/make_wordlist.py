#!/usr/bin/env python3
"""Dev-container-only fixture generator (reads /root/reference, which does not
exist on the GPU box).  Extracts word-type frequency tables from the reference's
own test corpora and writes them as small committed fixtures:

  tests/golden/en_wordlist.tsv   word<TAB>count   from data/botchan.txt
  tests/golden/ja_charlist.tsv   char<TAB>count   from data/wagahaiwa_nekodearu.txt

tools/corpus_gen.c turns these tables into the seeded synthetic corpora that
BASELINE.json's configs 2-5 name.
"""
import collections
import re
import sys

REF = "/root/reference/data"


def main():
    words = collections.Counter()
    with open(f"{REF}/botchan.txt", encoding="utf-8") as f:
        for line in f:
            for w in re.findall(r"[A-Za-z][a-z']*", line):
                if len(w) <= 24:
                    words[w] += 1
    with open("tests/golden/en_wordlist.tsv", "w", encoding="utf-8") as f:
        for w, c in sorted(words.items(), key=lambda kv: (-kv[1], kv[0])):
            f.write(f"{w}\t{c}\n")
    chars = collections.Counter()
    with open(f"{REF}/wagahaiwa_nekodearu.txt", encoding="utf-8") as f:
        for line in f:
            for ch in line:
                if ord(ch) >= 0x3000 and not ch.isspace():
                    chars[ch] += 1
    with open("tests/golden/ja_charlist.tsv", "w", encoding="utf-8") as f:
        for ch, c in sorted(chars.items(), key=lambda kv: (-kv[1], kv[0])):
            if c >= 2:
                f.write(f"{ch}\t{c}\n")
    print(len(words), "english word types;", sum(1 for c in chars.values() if c >= 2), "japanese chars")


if __name__ == "__main__":
    sys.exit(main())

"""Pins the oracle (oracle/spm_oracle.c) against the reference's OWN known-answer vectors
for the hot path, restated here from the reference's test sources (file:line cited per
block; paths relative to /root/reference).  CPU only."""
import os

import pytest

from conftest import ROOT, model_bytes
from oracle import modelproto as mp
from oracle import oracle_py

WS = "▁"
RC = "�"


def nmt_nfkc_charsmap():
    # the committed models were trained with the default normalization_rule_name=nmt_nfkc,
    # so their NormalizerSpec carries the nmt_nfkc blob the reference tests obtain from
    # SentencePieceTrainer::GetNormalizerSpec("nmt_nfkc") (normalizer_test.cc:32-34)
    return mp.parse_model(model_bytes("uni32k"))["charsmap"]


BASE = [("<unk>", 0.0, mp.UNKNOWN), ("<s>", 0.0, mp.CONTROL), ("</s>", 0.0, mp.CONTROL)]


def normalizer_model(**flags):
    flags.setdefault("charsmap", nmt_nfkc_charsmap())
    return oracle_py.OracleModel(mp.build_model(BASE + [("a", 0.0, mp.NORMAL)], **flags))


def norm(m, s):
    return m.normalize(s.encode("utf-8") if isinstance(s, str) else s)[0].decode("utf-8")


# src/normalizer_test.cc:37-75
def test_normalize_default():
    m = normalizer_model()
    assert norm(m, "") == ""
    assert norm(m, "      ") == ""
    assert norm(m, "　") == ""
    assert norm(m, "ABC") == WS + "ABC"
    assert norm(m, " ABC ") == WS + "ABC"
    assert norm(m, " A  B  C ") == WS + "A" + WS + "B" + WS + "C"
    assert norm(m, "   ABC   ") == WS + "ABC"
    assert norm(m, "   ＡＢＣ   ") == WS + "ABC"
    assert norm(m, "　　ABC") == WS + "ABC"
    assert norm(m, "　　ABC　　") == WS + "ABC"
    assert norm(m, "①②③") == WS + "123"
    assert norm(m, "㍿") == WS + "株式会社"
    assert norm(m, " ｸﾞｰｸﾞﾙ ") == WS + "グーグル"
    assert norm(m, " I  saw a　 　girl　　") == WS + "I" + WS + "saw" + WS + "a" + WS + "girl"
    for c in [0x7F, 0x8F, 0x9F, 0x0B] + list(range(0x10, 0x20)):
        assert norm(m, chr(c)) == ""


# src/normalizer_test.cc:77-95
def test_normalize_without_dummy_prefix():
    m = normalizer_model(add_dummy_prefix=False)
    assert norm(m, "") == "" and norm(m, "      ") == "" and norm(m, "　") == ""
    assert norm(m, "ABC") == "ABC"
    assert norm(m, " ABC ") == "ABC"
    assert norm(m, " A  B  C ") == "A" + WS + "B" + WS + "C"
    assert norm(m, "   ＡＢＣ   ") == "ABC"
    assert norm(m, "　　ABC　　") == "ABC"


# src/normalizer_test.cc:97-111
def test_normalize_ws_as_suffix():
    m = normalizer_model(treat_whitespace_as_suffix=True)
    assert norm(m, "") == "" and norm(m, "      ") == "" and norm(m, "　") == ""
    assert norm(m, "ABC") == "ABC" + WS
    assert norm(m, " ABC ") == "ABC" + WS
    assert norm(m, " A  B  C ") == "A" + WS + "B" + WS + "C" + WS
    assert norm(m, "   ABC   ") == "ABC" + WS


# src/normalizer_test.cc:113-128
def test_normalize_without_remove_extra_ws():
    m = normalizer_model(remove_extra_whitespaces=False)
    assert norm(m, "") == ""
    assert norm(m, "      ") == WS * 7
    assert norm(m, "　") == WS * 2
    assert norm(m, "ABC") == WS + "ABC"
    assert norm(m, " ABC ") == WS * 2 + "ABC" + WS
    assert norm(m, "  A  B  C  ") == WS * 3 + "A" + WS * 2 + "B" + WS * 2 + "C" + WS * 2


# src/normalizer_test.cc:130-147
def test_normalize_without_escape():
    m = normalizer_model(add_dummy_prefix=False, remove_extra_whitespaces=True, escape_whitespaces=False)
    assert norm(m, "") == "" and norm(m, "      ") == "" and norm(m, "　") == ""
    assert norm(m, "ABC") == "ABC"
    assert norm(m, " ABC ") == "ABC"
    assert norm(m, "  A  B  C  ") == "A B C"
    assert norm(m, "A　 B　 C") == "A B C"


def space_rules_blob():
    with open(os.path.join(ROOT, "tests", "golden", "charsmap_space_rules.bin"), "rb") as f:
        return f.read()


# src/normalizer_test.cc:149-264 (rules "a"->" A", "b"->"B", "c"->"D E", "d"->" F G ")
def test_space_contained_rules():
    blob = space_rules_blob()
    ins = ["a", "ba", "c", "da", "ad", "adb"]
    m = normalizer_model(charsmap=blob)
    assert [norm(m, s) for s in ins] == [WS + "A", WS + "B" + WS + "A", WS + "D" + WS + "E",
                                         WS + "F" + WS + "G" + WS + "A", WS + "A" + WS + "F" + WS + "G",
                                         WS + "A" + WS + "F" + WS + "G" + WS + "B"]
    table = [
        (dict(escape_whitespaces=False, add_dummy_prefix=False, remove_extra_whitespaces=True),
         ["A", "B A", "D E", "F G A", "A F G", "A F G B"]),
        (dict(escape_whitespaces=False, add_dummy_prefix=False, remove_extra_whitespaces=False),
         [" A", "B A", "D E", " F G  A", " A F G ", " A F G B"]),
        (dict(escape_whitespaces=False, add_dummy_prefix=True, remove_extra_whitespaces=True),
         [" A", " B A", " D E", " F G A", " A F G", " A F G B"]),
        (dict(escape_whitespaces=False, add_dummy_prefix=True, remove_extra_whitespaces=False),
         ["  A", " B A", " D E", "  F G  A", "  A F G ", "  A F G B"]),
    ]
    for flags, expected in table:
        m = normalizer_model(charsmap=blob, **flags)
        assert [norm(m, s) for s in ins] == expected, flags
    # kSpacePatternData, normalizer_test.cc:247-255: (dummy, remove_extra, escape, input, expected)
    pat = [(0, 0, 0, WS, WS), (0, 0, 1, WS, WS), (0, 1, 0, WS, WS), (0, 1, 1, WS, ""),
           (1, 0, 0, WS, " " + WS), (1, 0, 1, WS, WS + WS), (1, 1, 0, WS, " " + WS), (1, 1, 1, WS, ""),
           (0, 0, 0, " ", " "), (0, 0, 1, " ", WS), (0, 1, 0, " ", ""), (0, 1, 1, " ", ""),
           (1, 0, 0, " ", "  "), (1, 0, 1, " ", WS + WS), (1, 1, 0, " ", ""), (1, 1, 1, " ", "")]
    for d, r, e, i, exp in pat:
        m = normalizer_model(charsmap=blob, add_dummy_prefix=bool(d), remove_extra_whitespaces=bool(r),
                             escape_whitespaces=bool(e))
        assert norm(m, i) == exp, (d, r, e, i)


# src/normalizer_test.cc:266-275: malformed UTF-8 -> one U+FFFD per consumed byte
def test_replacement_char():
    m = normalizer_model(add_dummy_prefix=False)
    assert norm(m, b"abc\x80xy") == "abc" + RC + "xy"
    assert norm(m, b"abc\xc3") == "abc" + RC
    assert norm(m, b"ab\xe3\x81xy") == "ab" + RC + RC + "xy"
    assert norm(m, b"a\xf3\x81\x81xy") == "a" + RC * 3 + "xy"
    assert norm(m, b"ab\xc0\x82xy") == "ab" + RC + RC + "xy"


# src/normalizer_test.cc:277-357: norm_to_orig alignments
def test_norm_to_orig():
    m = normalizer_model()
    out, n2o = m.normalize("I saw a girl".encode())
    assert out.decode() == WS + "I" + WS + "saw" + WS + "a" + WS + "girl"
    assert n2o == [0, 0, 0, 0, 1, 1, 1, 2, 3, 4, 5, 5, 5, 6, 7, 7, 7, 8, 9, 10, 11, 12]
    out, n2o = m.normalize(" I   saw a　 　girl　　".encode())
    assert out.decode() == WS + "I" + WS + "saw" + WS + "a" + WS + "girl"
    assert n2o == [1, 1, 1, 1, 2, 2, 2, 5, 6, 7, 8, 8, 8, 9, 10, 10, 10, 17, 18, 19, 20, 21]
    out, n2o = m.normalize(" ｸﾞｰｸﾞﾙ ".encode())
    assert out.decode() == WS + "グーグル"
    assert n2o == [1, 1, 1, 1, 1, 1, 7, 7, 7, 10, 10, 10, 16, 16, 16, 19]
    out, n2o = m.normalize("①②③".encode())
    assert n2o == [0, 0, 0, 0, 3, 6, 9]
    out, n2o = m.normalize("㍿".encode())
    assert out.decode() == WS + "株式会社"
    assert n2o == [0] * 15 + [3]


def pieces_of(m, text):
    b = text.encode("utf-8") if isinstance(text, str) else text
    ids, ends = m.model_encode(b)
    out, prev = [], 0
    for e in ends:
        out.append(b[prev:int(e)])
        prev = int(e)
    return [p.decode("utf-8", "surrogateescape") for p in out], [int(i) for i in ids]


ENCODE_PIECES = [("ab", 0.0), ("cd", -0.1), ("abc", -0.2), ("a", -0.3), ("b", -0.4), ("c", -0.5), ("ABC", -0.5),
                 ("abcdabcd", -0.5), ("q", -0.5), ("r", -0.5), ("qr", -0.5)]


def encode_test_model(model_type):
    pcs = BASE + [(p, s, mp.NORMAL) for p, s in ENCODE_PIECES]
    for i in (9, 10, 11, 12):  # ABC, abcdabcd, q, r are USER_DEFINED
        pcs[i] = (pcs[i][0], pcs[i][1], mp.USER_DEFINED)
    return oracle_py.OracleModel(mp.build_model(pcs, model_type=model_type, charsmap=b""))


# src/unigram_model_test.cc:782-871
def test_unigram_encode():
    m = encode_test_model(mp.UNIGRAM)
    assert pieces_of(m, "abc")[0] == ["abc"]
    assert pieces_of(m, "AB")[0] == ["A", "B"]
    assert pieces_of(m, "abcd")[0] == ["ab", "cd"]
    assert pieces_of(m, "abcc")[0] == ["abc", "c"]
    assert pieces_of(m, "xabcabaabcdd")[0] == ["x", "abc", "ab", "a", "ab", "cd", "d"]
    assert pieces_of(m, "xyz東京")[0] == ["x", "y", "z", "東", "京"]
    assert pieces_of(m, "ABC")[0] == ["ABC"]
    assert pieces_of(m, "abABCcd")[0] == ["ab", "ABC", "cd"]
    assert pieces_of(m, "ababcdabcdcd")[0] == ["ab", "abcdabcd", "cd"]
    assert pieces_of(m, "abqrcd")[0] == ["ab", "q", "r", "cd"]


UNUSED_PIECES = [("abcd", 10.0), ("abc", 5.0), ("ab", 2.0), ("cd", 1.0), ("a", 0.0), ("b", 0.0), ("c", 0.0), ("d", 0.0)]


def unused_model(model_type, unused_ids):
    pcs = BASE + [(p, s, mp.UNUSED if 3 + i in unused_ids else mp.NORMAL) for i, (p, s) in enumerate(UNUSED_PIECES)]
    return oracle_py.OracleModel(mp.build_model(pcs, model_type=model_type, charsmap=b""))


# src/unigram_model_test.cc:873-928
def test_unigram_unused():
    assert pieces_of(unused_model(mp.UNIGRAM, ()), "abcd")[0] == ["abcd"]
    assert pieces_of(unused_model(mp.UNIGRAM, (3,)), "abcd")[0] == ["abc", "d"]
    assert pieces_of(unused_model(mp.UNIGRAM, (3, 5)), "abcd")[0] == ["abc", "d"]
    assert pieces_of(unused_model(mp.UNIGRAM, (3, 4)), "abcd")[0] == ["ab", "cd"]


# src/bpe_model_test.cc:195-250
def test_bpe_unused():
    assert pieces_of(unused_model(mp.BPE, ()), "abcd")[0] == ["abcd"]
    assert pieces_of(unused_model(mp.BPE, (3,)), "abcd")[0] == ["abc", "d"]
    assert pieces_of(unused_model(mp.BPE, (3, 5)), "abcd")[0] == ["abc", "d"]
    assert pieces_of(unused_model(mp.BPE, (3, 4)), "abcd")[0] == ["ab", "c", "d"]


# src/bpe_model_test.cc:143-187: ties resolve leftmost first; broken UTF-8 is one symbol
def test_bpe_ambiguous():
    pcs = BASE + [(p, s, mp.NORMAL) for p, s in [("aa", -0.1), ("bb", -0.2), ("ab", -0.3), ("a", -0.4), ("b", -0.5)]]
    m = oracle_py.OracleModel(mp.build_model(pcs, model_type=mp.BPE, charsmap=b""))
    assert pieces_of(m, "aaa")[0] == ["aa", "a"]
    assert pieces_of(m, "aabb")[0] == ["aa", "bb"]
    assert pieces_of(m, "aaabbb")[0] == ["aa", "a", "bb", "b"]
    assert pieces_of(m, "aaaba")[0] == ["aa", "ab", "a"]
    ids, ends = m.model_encode(b"\xe3")
    assert list(ends) == [1]


# src/bpe_model_test.cc:49-141 (same inputs as the unigram EncodeTest, BPE semantics)
def test_bpe_encode():
    m = encode_test_model(mp.BPE)
    assert pieces_of(m, "abc")[0] == ["abc"]
    assert pieces_of(m, "AB")[0] == ["A", "B"]
    assert pieces_of(m, "abcd")[0] == ["ab", "cd"]
    assert pieces_of(m, "abcc")[0] == ["abc", "c"]
    assert pieces_of(m, "xabcabaabcdd")[0] == ["x", "abc", "ab", "a", "ab", "cd", "d"]
    assert pieces_of(m, "xyz東京")[0] == ["x", "y", "z", "東", "京"]
    assert pieces_of(m, "ABC")[0] == ["ABC"]
    assert pieces_of(m, "abABCcd")[0] == ["ab", "ABC", "cd"]
    assert pieces_of(m, "ababcdabcdcd")[0] == ["ab", "abcdabcd", "cd"]
    assert pieces_of(m, "abqrcd")[0] == ["ab", "q", "r", "cd"]


# src/sentencepiece_processor_test.cc:186-233: unknown runs merge into one piece / one id
def test_unk_run_merging():
    pcs = BASE + [("▁", -1.0, mp.NORMAL), ("a", -1.0, mp.NORMAL), ("b", -1.0, mp.NORMAL)]
    m = oracle_py.OracleModel(mp.build_model(pcs, charsmap=b""))
    ids, te = m.encode("axyzb".encode())
    assert list(ids) == [3, 4, 0, 5]
    assert list(te) == [3, 4, 7, 8]  # "▁" "a" "xyz"(merged) "b" in the normalized text "▁axyzb"


# src/sentencepiece_processor_test.cc:235-303: byte fallback expands every byte of an unknown piece
def test_byte_fallback_expansion():
    pcs = BASE + [("<0x%02X>" % b, 0.0, mp.BYTE) for b in range(256)]
    pcs += [("▁", -1.0, mp.NORMAL), ("a", -1.0, mp.NORMAL)]
    m = oracle_py.OracleModel(mp.build_model(pcs, byte_fallback=True, charsmap=b""))
    ids, te = m.encode("aあ".encode())
    assert list(ids) == [3 + 256, 3 + 257, 3 + 0xE3, 3 + 0x81, 3 + 0x82]
    assert list(te) == [3, 4, 5, 6, 7]


# src/util_test.cc:127-226 (DecodeUTF8 validity rules) observed through the normalizer:
# invalid sequences consume one byte and become U+FFFD; a literal U+FFFD stays.
def test_utf8_validity_rules():
    m = normalizer_model(add_dummy_prefix=False, charsmap=b"")
    assert norm(m, b"\xc0\xaf") == RC * 2          # overlong 2-byte
    assert norm(m, b"\xe0\x80\xaf") == RC * 3      # overlong 3-byte
    assert norm(m, b"\xed\xa0\x80") == RC * 3      # surrogate
    assert norm(m, b"\xf4\x90\x80\x80") == RC * 4  # > U+10FFFF
    assert norm(m, b"\xf8\x88\x80\x80\x80") == RC * 5
    assert norm(m, "�".encode()) == RC        # literal replacement char is valid
    assert norm(m, "\U0010ffff".encode()) == "\U0010ffff"
    assert norm(m, b"\xe3\x81") == RC * 2


# unigram_model.cc:657-664: max_score_ starts at FLT_MIN (quirk Q3); unk score = min - 10
def test_score_quirks():
    m = oracle_py.OracleModel(model_bytes("uni32k"))
    assert m.max_score == pytest.approx(1.1754943508222875e-38, rel=0, abs=0)
    assert m.min_score < -10


# models trained with --self_test_sample_size embed (input, expected pieces) pairs that the
# reference re-verifies on Load (sentencepiece_processor.cc:259-278): a built-in KAT.
def test_embedded_self_test():
    mb = model_bytes("botchan8k")
    proto = mp.parse_model(mb)
    assert proto["self_test"], "botchan8k was trained with --self_test_sample_size=20"
    m = oracle_py.OracleModel(mb)
    for inp, expected in proto["self_test"]:
        ids, te = m.encode(inp)
        normd = m.normalize(inp)[0]
        prev, got = 0, []
        for i, e in zip(ids, te):
            got.append(normd[prev:int(e)] if int(i) == m.unk_id else proto["pieces"][int(i)])
            prev = int(e)
        assert b" ".join(got) == expected

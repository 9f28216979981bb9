"""One-off GPU check: adversarial byte mixes (the generator of tests/test_oracle_golden.py::test_fuzz_oracle_vs_live_reference)
through the engine's encode and decode, compared with the oracle.  usage: python tools/gpu_fuzz.py [n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py  # noqa: E402
from sentencepiece_b200 import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(777)
chunks = [b" ", b"  ", b"a", b"e", b"the", b"ing", "▁".encode(), "あ".encode(), "ガ".encode(), "ｗ".encode(),
          "㍿".encode(), "😀".encode(), b"\xff", b"\xc0\xaf", b"\xed\xa0\x80", b"\xe2\x82", b"\x00", b"\t", b"\n",
          "Å".encode(), b"1", "①".encode(), b".", b",", " ".encode(), "　".encode(), b"<unk>", b"<s>", "�".encode(),
          b"word ", b" and", b"ab ", b"abc", b"xyzA", "́".encode()]
sents = []
for _ in range(n):
    parts = [chunks[int(rng.integers(0, len(chunks)))] for _ in range(int(rng.integers(0, 40)))]
    if rng.random() < 0.2:
        parts.append(bytes(rng.integers(0, 256, size=int(rng.integers(1, 12)), dtype=np.uint8)))
    sents.append(b"".join(parts))
buf, offs = oracle_py.pack(sents)
bad = 0
for model in ("uni32k", "mix_bf8k", "bpe32k"):
    mb = open(os.path.join(ROOT, "tests", "golden", "models", model + ".model"), "rb").read()
    eng = Engine(mb)
    om = oracle_py.OracleModel(mb)
    ids, ido = eng.encode_packed(buf, offs)
    oids, oido = om.encode_batch(buf, offs)
    ok = np.array_equal(ido, oido) and np.array_equal(ids, oids)
    text, to = eng.decode_packed(oids, oido)
    otext, oto = om.decode_batch(oids, oido)
    okd = np.array_equal(to, oto) and np.array_equal(text, otext)
    print(model, "encode", "OK" if ok else "MISMATCH", "decode", "OK" if okd else "MISMATCH", len(oids), "ids")
    if not ok:
        for i in range(n):
            x, y = ids[int(ido[i]):int(ido[i + 1])], oids[int(oido[i]):int(oido[i + 1])]
            if len(x) != len(y) or (x != y).any():
                print("  first differing sentence", i, sents[i])
                break
    bad += (not ok) + (not okd)
    eng.close()
sys.exit(1 if bad else 0)

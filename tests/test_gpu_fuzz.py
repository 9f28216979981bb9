"""Adversarial byte mixes through the engine (encode and decode) against the oracle: ASCII words and runs of spaces
next to CJK, emoji, NFKC compatibility forms, combining marks, control bytes, NUL, reserved piece strings and malformed
UTF-8 -- the inputs that decide between the normalizer's 4-byte step and its byte-by-byte path.  Needs a B200.
(Same generator as tools/gpu_fuzz.py.)"""
import numpy as np
import pytest

from conftest import model_bytes
from oracle import oracle_py

pytestmark = pytest.mark.gpu


def _sentences(n, seed):
    rng = np.random.default_rng(seed)
    chunks = [b" ", b"  ", b"a", b"e", b"the", b"ing", "▁".encode(), "あ".encode(), "ガ".encode(), "ｗ".encode(),
              "㍿".encode(), "😀".encode(), b"\xff", b"\xc0\xaf", b"\xed\xa0\x80", b"\xe2\x82", b"\x00", b"\t", b"\n",
              "Å".encode(), b"1", "①".encode(), b".", b",", " ".encode(), "　".encode(), b"<unk>", b"<s>",
              "�".encode(), b"word ", b" and", b"ab ", b"abc", b"xyzA", "́".encode()]
    sents = []
    for _ in range(n):
        parts = [chunks[int(rng.integers(0, len(chunks)))] for _ in range(int(rng.integers(0, 40)))]
        if rng.random() < 0.2:
            parts.append(bytes(rng.integers(0, 256, size=int(rng.integers(1, 12)), dtype=np.uint8)))
        sents.append(b"".join(parts))
    return sents


@pytest.mark.parametrize("model", ["uni32k", "mix_bf8k", "bpe32k"])
def test_fuzz_encode_decode(model):
    from sentencepiece_b200 import Engine
    buf, offs = oracle_py.pack(_sentences(20000, 777))
    mb = model_bytes(model)
    eng = Engine(mb)
    om = oracle_py.OracleModel(mb)
    ids, ido = eng.encode_packed(buf, offs)
    oids, oido = om.encode_batch(buf, offs)
    assert np.array_equal(ido, oido) and np.array_equal(ids, oids)
    text, to = eng.decode_packed(oids, oido)
    otext, oto = om.decode_batch(oids, oido)
    assert np.array_equal(to, oto) and np.array_equal(text, otext)
    eng.close()

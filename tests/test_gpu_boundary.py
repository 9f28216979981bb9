"""The drop-in boundary as INTEGRATION.md shows it to a maintainer: an engine built from the flattened
`spm_model_desc` (what SentencePieceProcessor::Load would fill from model_proto_) must behave exactly like one built
from the serialized ModelProto, and `spm_engine_set_unk_surface` must carry TrainerSpec.unk_surface into Decode.
Also: a synthetic model with very long pieces (the trainer's default of 16 characters per piece allows ~50-byte CJK
pieces) must still encode -- the lane kernel shrinks its CTA until the rings fit, or the tile kernel takes over.
Needs a B200."""
import ctypes

import numpy as np
import pytest

from conftest import model_bytes
from oracle import modelproto as mp
from oracle import oracle_py

pytestmark = pytest.mark.gpu


class DescEngine:
    """An engine created through spm_engine_create(&spm_model_desc) (the Engine class uses _from_serialized)."""

    def __init__(self, mbytes, device=0):
        from sentencepiece_b200 import Engine, _capi
        self.lib = _capi.load()
        m = mp.parse_model(mbytes)
        blob = b"".join(m["pieces"])
        off = np.zeros(len(m["pieces"]) + 1, dtype=np.uint32)
        off[1:] = np.cumsum([len(p) for p in m["pieces"]])
        scores = np.asarray(m["scores"], dtype=np.float32)
        types = np.asarray(m["types"], dtype=np.uint8)
        d = _capi.ModelDesc(m["model_type"], len(m["pieces"]), blob, off.ctypes.data, scores.ctypes.data, types.ctypes.data,
                            int(m["byte_fallback"]), int(m["treat_whitespace_as_suffix"]), int(m["add_dummy_prefix"]),
                            int(m["remove_extra_whitespaces"]), int(m["escape_whitespaces"]), (ctypes.c_uint8 * 3)(),
                            m["charsmap"], len(m["charsmap"]))
        h = ctypes.c_void_p()
        rc = self.lib.spm_engine_create(ctypes.byref(d), device, ctypes.byref(h))
        assert rc == 0, self.lib.spm_last_error(None).decode()
        self.eng = Engine.__new__(Engine)   # borrow the packed-batch helpers of the ctypes mirror
        self.eng._lib = self.lib
        self.eng._h = h
        self.proto = m

    def close(self):
        self.eng.close()


@pytest.mark.parametrize("model,kind", [("uni32k", "en"), ("mix_bf8k", "mixed"), ("bpe32k", "en"), ("mix_bpe4k", "mixed"),
                                        ("botchan8k", "mixed")])
def test_engine_from_model_desc(model, kind, corpus_gen):
    from sentencepiece_b200 import Engine
    mb = model_bytes(model)
    lines = corpus_gen.lines(kind, 7301, 5000) + [b"", b"   ", "\U0001F600 unknown あい".encode()]
    buf, offs = oracle_py.pack(lines)
    de = DescEngine(mb)
    se = Engine(mb)
    ids_d, ido_d = de.eng.encode_packed(buf, offs)
    ids_s, ido_s = se.encode_packed(buf, offs)
    om = oracle_py.OracleModel(mb)
    oids, oido = om.encode_batch(buf, offs)
    assert np.array_equal(ido_d, ido_s) and np.array_equal(ids_d, ids_s)
    assert np.array_equal(ido_d, oido) and np.array_equal(ids_d, oids)
    # spans (pieces / alignment) go through the same tables
    sd, ss = de.eng.encode_spans(buf, offs), se.encode_spans(buf, offs)
    for k in ("ids", "tok_end", "id_offsets", "norm_offsets", "n2o"):
        assert np.array_equal(sd[k], ss[k]), k
    assert sd["normalized"] == ss["normalized"]
    # Decode: the desc carries no TrainerSpec.unk_surface -- set it like a maintainer would, then compare with the
    # serialized engine (which read it from the proto) and the oracle, including a non-default surface
    for surface in (de.proto["unk_surface"], b"<??>", b""):
        assert de.lib.spm_engine_set_unk_surface(de.eng._h, surface, len(surface)) == 0
        assert se._lib.spm_engine_set_unk_surface(se._h, surface, len(surface)) == 0
        om.lib.oracle_set_unk_surface(om.h, surface, len(surface))
        td, tod = de.eng.decode_packed(ids_d, ido_d)
        ts, tos = se.decode_packed(ids_d, ido_d)
        ot, oto = om.decode_batch(ids_d, ido_d)
        assert np.array_equal(tod, tos) and np.array_equal(td, ts)
        assert np.array_equal(tod, oto) and np.array_equal(td, ot)
    de.close()
    se.close()


def _long_piece_model(n_long, byte_len_chars):
    """A unigram model whose longest pieces are `byte_len_chars` three-byte characters long."""
    base = [("<unk>", 0.0, mp.UNKNOWN), ("<s>", 0.0, mp.CONTROL), ("</s>", 0.0, mp.CONTROL), ("▁", -2.0, mp.NORMAL)]
    chars = [chr(0x4E00 + i) for i in range(40)]
    pcs = base + [(c, -6.0 - 0.01 * i, mp.NORMAL) for i, c in enumerate(chars)]
    pcs += [(c, -5.0 - 0.01 * i, mp.NORMAL) for i, c in enumerate("abcdefghijklmnopqrstuvwxyz")]
    rng = np.random.RandomState(3)
    seen = set()
    for k in range(n_long):
        ln = int(rng.randint(2, byte_len_chars + 1))
        s = "".join(chars[int(x)] for x in rng.randint(0, len(chars), ln))
        if s in seen:
            continue
        seen.add(s)
        pcs.append((s, -8.0 - 0.001 * k - 0.3 * ln, mp.NORMAL))
    longest = "".join(chars[i % len(chars)] for i in range(byte_len_chars))
    if longest not in seen:
        pcs.append((longest, -9.0, mp.NORMAL))
    return mp.build_model(pcs, charsmap=b"")


@pytest.mark.parametrize("chars_per_piece", [9, 16, 20, 30])  # 27, 48, 60 and 90 byte pieces
def test_long_pieces(chars_per_piece):
    """with 26-62 byte pieces the lane kernel's rings did not fit 32 warps and every ids-only
    encode failed; > 62 bytes is outside the lane kernels altogether (tile kernel)."""
    from sentencepiece_b200 import Engine
    mb = _long_piece_model(400, chars_per_piece)
    m = mp.parse_model(mb)
    rng = np.random.RandomState(11)
    cjk = [p.decode() for p in m["pieces"][4:]]
    lines = []
    for _ in range(3000):
        k = int(rng.randint(1, 12))
        lines.append(" ".join(cjk[int(x)] for x in rng.randint(0, len(cjk), k)).encode())
    buf, offs = oracle_py.pack(lines)
    eng = Engine(mb)
    ids, ido = eng.encode_packed(buf, offs)
    oids, oido = oracle_py.OracleModel(mb).encode_batch(buf, offs)
    assert np.array_equal(ido, oido) and np.array_equal(ids, oids)
    eng.close()

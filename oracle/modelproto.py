"""oracle/modelproto.py -- TEST INFRASTRUCTURE ONLY.

A tiny, independent protobuf wire-format reader/writer for the reference's
ModelProto (src/sentencepiece_model.proto:24-332).  The product has its own C++
reader (sentencepiece_b200/csrc/model_reader.cc); keeping this one separate means
the oracle and the engine never share model-parsing code, so a parsing bug in
either shows up as a parity failure.
"""
import struct

# SentencePiece.Type (sentencepiece_model.proto:296-304)
NORMAL, UNKNOWN, CONTROL, USER_DEFINED, UNUSED, BYTE = 1, 2, 3, 4, 5, 6
UNIGRAM, BPE, WORD, CHAR = 1, 2, 3, 4


def _varint(buf, pos):
    shift = 0
    val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _fields(buf):
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def parse_model(data):
    """Returns a dict with the fields the encode path reads."""
    data = bytes(data)
    m = dict(pieces=[], scores=[], types=[], model_type=UNIGRAM, byte_fallback=False,
             treat_whitespace_as_suffix=False, unk_piece=b"<unk>", charsmap=b"",
             add_dummy_prefix=True, remove_extra_whitespaces=True, escape_whitespaces=True,
             self_test=[], has_normalizer_spec=False, unk_surface=" \u2047 ".encode(), denormalizer_charsmap=b"")
    for fno, wt, v in _fields(data):
        if fno == 1 and wt == 2:  # SentencePiece
            piece, score, typ = b"", 0.0, NORMAL
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    piece = bytes(v2)
                elif f2 == 2:
                    score = struct.unpack("<f", v2)[0]
                elif f2 == 3:
                    typ = v2
            m["pieces"].append(piece)
            m["scores"].append(score)
            m["types"].append(typ)
        elif fno == 2 and wt == 2:  # TrainerSpec
            for f2, w2, v2 in _fields(v):
                if f2 == 3:
                    m["model_type"] = v2
                elif f2 == 35:
                    m["byte_fallback"] = bool(v2)
                elif f2 == 24:
                    m["treat_whitespace_as_suffix"] = bool(v2)
                elif f2 == 45:
                    m["unk_piece"] = bytes(v2)
                elif f2 == 44:
                    m["unk_surface"] = bytes(v2)
        elif fno == 3 and wt == 2:  # NormalizerSpec
            m["has_normalizer_spec"] = True
            for f2, w2, v2 in _fields(v):
                if f2 == 2:
                    m["charsmap"] = bytes(v2)
                elif f2 == 3:
                    m["add_dummy_prefix"] = bool(v2)
                elif f2 == 4:
                    m["remove_extra_whitespaces"] = bool(v2)
                elif f2 == 5:
                    m["escape_whitespaces"] = bool(v2)
        elif fno == 5 and wt == 2:  # denormalizer_spec (NormalizerSpec)
            for f2, w2, v2 in _fields(v):
                if f2 == 2:
                    m["denormalizer_charsmap"] = bytes(v2)
        elif fno == 4 and wt == 2:  # SelfTestData
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    inp = exp = b""
                    for f3, w3, v3 in _fields(v2):
                        if f3 == 1:
                            inp = bytes(v3)
                        elif f3 == 2:
                            exp = bytes(v3)
                    m["self_test"].append((inp, exp))
    return m


# ------------------------------------------------------------------ writer --

def _enc_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(fno, wt):
    return _enc_varint((fno << 3) | wt)


def _ld(fno, payload):
    return _key(fno, 2) + _enc_varint(len(payload)) + payload


def build_model(pieces, model_type=UNIGRAM, byte_fallback=False, treat_whitespace_as_suffix=False,
                charsmap=b"", add_dummy_prefix=True, remove_extra_whitespaces=True, escape_whitespaces=True,
                with_normalizer_spec=True):
    """pieces: list of (piece: bytes|str, score: float, type: int).  Mirrors what the
    reference's tests assemble with MakeBaseModelProto/AddPiece
    (src/unigram_model_test.cc:468-504, src/bpe_model_test.cc:26-47)."""
    out = bytearray()
    for p, s, t in pieces:
        if isinstance(p, str):
            p = p.encode("utf-8")
        sp = _ld(1, p) + _key(2, 5) + struct.pack("<f", s)
        if t != NORMAL:
            sp += _key(3, 0) + _enc_varint(t)
        out += _ld(1, sp)
    ts = _key(3, 0) + _enc_varint(model_type)
    if treat_whitespace_as_suffix:
        ts += _key(24, 0) + _enc_varint(1)
    if byte_fallback:
        ts += _key(35, 0) + _enc_varint(1)
    out += _ld(2, ts)
    if with_normalizer_spec:
        ns = b""
        if charsmap:
            ns += _ld(2, charsmap)
        ns += _key(3, 0) + _enc_varint(int(add_dummy_prefix))
        ns += _key(4, 0) + _enc_varint(int(remove_extra_whitespaces))
        ns += _key(5, 0) + _enc_varint(int(escape_whitespaces))
        out += _ld(3, ns)
    return bytes(out)


def replace_flags(data, **kw):
    """Re-serialise `data` with normalizer/trainer flags overridden (used by the
    flag-variant KATs, src/normalizer_test.cc:77-147)."""
    m = parse_model(data)
    args = dict(model_type=m["model_type"], byte_fallback=m["byte_fallback"],
                treat_whitespace_as_suffix=m["treat_whitespace_as_suffix"], charsmap=m["charsmap"],
                add_dummy_prefix=m["add_dummy_prefix"], remove_extra_whitespaces=m["remove_extra_whitespaces"],
                escape_whitespaces=m["escape_whitespaces"])
    pieces = kw.pop("pieces", None)
    args.update(kw)
    if pieces is None:
        pieces = list(zip(m["pieces"], m["scores"], m["types"]))
    return build_model(pieces, **args)

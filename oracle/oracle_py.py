"""oracle/oracle_py.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-ends for the two checkers:
  OracleModel : oracle/spm_oracle.c  (our plain-C restatement)
  RefModel    : oracle/_ref/libspm_ref_shim.so (the UNMODIFIED reference, compiled by
                oracle/Makefile from /root/reference; prebuilt files travel to the GPU box)

May be imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import modelproto

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle_spm.so")
_REF_SO = os.path.join(_HERE, "_ref", "libspm_ref_shim.so")


def build_oracle(force=False):
    src = os.path.join(_HERE, "spm_oracle.c")
    if force or not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _ORACLE_SO


def ref_available():
    return os.path.exists(_REF_SO)


class _Desc(ctypes.Structure):
    _fields_ = [("model_type", ctypes.c_int32), ("vocab_size", ctypes.c_int32),
                ("piece_bytes", ctypes.c_char_p), ("piece_off", ctypes.c_void_p),
                ("scores", ctypes.c_void_p), ("types", ctypes.c_void_p),
                ("byte_fallback", ctypes.c_uint8), ("add_dummy_prefix", ctypes.c_uint8),
                ("remove_extra_whitespaces", ctypes.c_uint8), ("escape_whitespaces", ctypes.c_uint8),
                ("treat_whitespace_as_suffix", ctypes.c_uint8),
                ("charsmap", ctypes.c_char_p), ("charsmap_len", ctypes.c_size_t)]


def pack(sentences):
    """list[bytes] -> (uint8 array, uint64 offsets[n+1])"""
    offs = np.zeros(len(sentences) + 1, dtype=np.uint64)
    if sentences:
        offs[1:] = np.cumsum([len(s) for s in sentences], dtype=np.uint64)
    buf = np.frombuffer(b"".join(sentences), dtype=np.uint8) if sentences else np.zeros(0, np.uint8)
    return buf, offs


def _i32(ptr, k):
    if not k:
        return np.zeros(0, np.int32)
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int32)), (k,)).copy()


def _u32(ptr, k):
    if not k:
        return np.zeros(0, np.uint32)
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint32)), (k,)).copy()


def _u64(ptr, k):
    if not k:
        return np.zeros(0, np.uint64)
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint64)), (k,)).copy()


class OracleModel:
    def __init__(self, model_bytes):
        self.lib = ctypes.CDLL(build_oracle())
        L = self.lib
        L.oracle_create.restype = ctypes.c_void_p
        L.oracle_create.argtypes = [ctypes.POINTER(_Desc), ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_set_types.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_free.argtypes = [ctypes.c_void_p]
        L.oracle_min_score.restype = ctypes.c_float
        L.oracle_min_score.argtypes = [ctypes.c_void_p]
        L.oracle_max_score.restype = ctypes.c_float
        L.oracle_max_score.argtypes = [ctypes.c_void_p]
        L.oracle_unk_id.argtypes = [ctypes.c_void_p]
        L.oracle_normalize.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t,
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]
        L.oracle_encode.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t,
                                    ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                    ctypes.POINTER(ctypes.c_size_t)]
        L.oracle_model_encode.argtypes = L.oracle_encode.argtypes
        L.oracle_encode_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                          ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]
        self.proto = m = modelproto.parse_model(model_bytes)
        blob = b"".join(m["pieces"])
        off = np.zeros(len(m["pieces"]) + 1, dtype=np.uint32)
        if m["pieces"]:
            off[1:] = np.cumsum([len(p) for p in m["pieces"]])
        scores = np.asarray(m["scores"], dtype=np.float32)
        self.types = np.asarray(m["types"], dtype=np.uint8)
        d = _Desc(m["model_type"], len(m["pieces"]), blob, off.ctypes.data, scores.ctypes.data,
                  self.types.ctypes.data, int(m["byte_fallback"]), int(m["add_dummy_prefix"]),
                  int(m["remove_extra_whitespaces"]), int(m["escape_whitespaces"]),
                  int(m["treat_whitespace_as_suffix"]), m["charsmap"], len(m["charsmap"]))
        err = ctypes.create_string_buffer(256)
        self.h = L.oracle_create(ctypes.byref(d), err, 256)
        if not self.h:
            raise ValueError("oracle_create: " + err.value.decode())
        L.oracle_set_unk_surface.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_set_unk_surface(self.h, m["unk_surface"], len(m["unk_surface"]))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_destroy(self.h)
            self.h = None

    @property
    def min_score(self):
        return self.lib.oracle_min_score(self.h)

    @property
    def max_score(self):
        return self.lib.oracle_max_score(self.h)

    @property
    def unk_id(self):
        return self.lib.oracle_unk_id(self.h)

    def set_types(self, types):
        t = np.ascontiguousarray(types, dtype=np.uint8)
        self.lib.oracle_set_types(self.h, t.ctypes.data)

    def vocabulary_types(self, valid):
        """Piece types after SetVocabulary(valid), src/sentencepiece_processor.cc:301-330."""
        valid = {v.encode() if isinstance(v, str) else v for v in valid}
        t = self.types.copy()
        for i, p in enumerate(self.proto["pieces"]):
            if t[i] in (modelproto.CONTROL, modelproto.UNKNOWN, modelproto.USER_DEFINED):
                continue
            # pieces of one unicode character are always kept (:318-321)
            if p in valid or len(p.decode("utf-8", "replace")) == 1:
                t[i] = modelproto.NORMAL
            else:
                t[i] = modelproto.UNUSED
        return t

    def normalize(self, s):
        out = ctypes.c_void_p()
        n = ctypes.c_size_t()
        n2o = ctypes.c_void_p()
        n2on = ctypes.c_size_t()
        rc = self.lib.oracle_normalize(self.h, s, len(s), ctypes.byref(out), ctypes.byref(n), ctypes.byref(n2o),
                                       ctypes.byref(n2on))
        if rc:
            raise RuntimeError(f"oracle_normalize rc={rc}")
        res = ctypes.string_at(out, n.value) if n.value else b""
        m = [int(x) for x in _u64(n2o, n2on.value)]
        self.lib.oracle_free(out)
        self.lib.oracle_free(n2o)
        return res, m

    def _enc(self, fn, s):
        ids = ctypes.c_void_p()
        te = ctypes.c_void_p()
        n = ctypes.c_size_t()
        rc = fn(self.h, s, len(s), ctypes.byref(ids), ctypes.byref(te), ctypes.byref(n))
        if rc:
            raise RuntimeError(f"oracle encode rc={rc}")
        a, b = _i32(ids, n.value), _u32(te, n.value)
        self.lib.oracle_free(ids)
        self.lib.oracle_free(te)
        return a, b

    def encode(self, s):
        """full path: (ids int32[], tok_end uint32[])"""
        return self._enc(self.lib.oracle_encode, s)

    def model_encode(self, normalized):
        return self._enc(self.lib.oracle_model_encode, normalized)

    def nbest_encode(self, s, nbest_size):
        """-> (list of int32 id arrays, float32 scores) like SentencePieceProcessor::NBestEncode"""
        L = self.lib
        L.oracle_nbest_encode.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                          ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]
        ids, co, sc, k = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t()
        rc = L.oracle_nbest_encode(self.h, s, len(s), nbest_size, ctypes.byref(ids), ctypes.byref(co), ctypes.byref(sc),
                                   ctypes.byref(k))
        if rc:
            raise RuntimeError(f"oracle_nbest_encode rc={rc}")
        kk = k.value
        off = _u32(co, kk + 1)
        allids = _i32(ids, int(off[kk]))
        scores = np.ctypeslib.as_array(ctypes.cast(sc, ctypes.POINTER(ctypes.c_float)), (kk,)).copy()
        for ptr in (ids, co, sc):
            L.oracle_free(ptr)
        return [allids[int(off[i]):int(off[i + 1])] for i in range(kk)], scores

    def sample_encode_batch(self, buf, offs, nbest_size, alpha, seed):
        L = self.lib
        L.oracle_sample_encode_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                 ctypes.c_int, ctypes.c_float, ctypes.c_uint32,
                                                 ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]
        n = len(offs) - 1
        ido = np.zeros(n + 1, dtype=np.uint64)
        ids = ctypes.c_void_p()
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        rc = L.oracle_sample_encode_batch(self.h, buf.ctypes.data, offs.ctypes.data, n, nbest_size, alpha, seed,
                                          ctypes.byref(ids), ido.ctypes.data)
        if rc:
            raise RuntimeError(f"oracle_sample_encode_batch failed at sentence {rc - 1}")
        a = _i32(ids, int(ido[n]))
        L.oracle_free(ids)
        return a, ido

    def entropy_batch(self, buf, offs, alpha):
        """CalculateEntropy per sentence -> float32[n]"""
        L = self.lib
        L.oracle_entropy_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                           ctypes.c_float, ctypes.c_void_p]
        n = len(offs) - 1
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        ent = np.zeros(n, dtype=np.float32)
        rc = L.oracle_entropy_batch(self.h, buf.ctypes.data, offs.ctypes.data, n, alpha, ent.ctypes.data)
        if rc:
            raise RuntimeError(f"oracle_entropy_batch failed ({rc})")
        return ent

    def sample_score_batch(self, buf, offs, samples, alpha, seed):
        """SampleEncodeAndScore(wor=False) -> (ids, cand_off uint64[n*samples+1], scores float32[n*samples])"""
        L = self.lib
        L.oracle_sample_score_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                ctypes.c_int, ctypes.c_float, ctypes.c_uint32,
                                                ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_void_p]
        n = len(offs) - 1
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        co = np.zeros(n * samples + 1, dtype=np.uint64)
        sc = np.zeros(n * samples, dtype=np.float32)
        ids = ctypes.c_void_p()
        rc = L.oracle_sample_score_batch(self.h, buf.ctypes.data, offs.ctypes.data, n, samples, alpha, seed,
                                         ctypes.byref(ids), co.ctypes.data, sc.ctypes.data)
        if rc:
            raise RuntimeError(f"oracle_sample_score_batch failed ({rc})")
        a = _i32(ids, int(co[-1]))
        L.oracle_free(ids)
        return a, co, sc

    def encode_batch(self, buf, offs):
        n = len(offs) - 1
        ido = np.zeros(n + 1, dtype=np.uint64)
        ids = ctypes.c_void_p()
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        rc = self.lib.oracle_encode_batch(self.h, buf.ctypes.data, offs.ctypes.data, n, ctypes.byref(ids),
                                          ido.ctypes.data)
        if rc:
            raise RuntimeError(f"oracle_encode_batch failed at sentence {rc - 1}")
        a = _i32(ids, int(ido[n]))
        self.lib.oracle_free(ids)
        return a, ido


    def decode_batch(self, ids, ido):
        """Decode(ids) per list -> (text uint8[], text_offsets uint64[n+1])"""
        L = self.lib
        L.oracle_decode_ids.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p),
                                        ctypes.POINTER(ctypes.c_size_t)]
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        n = len(ido) - 1
        parts = []
        to = np.zeros(n + 1, dtype=np.uint64)
        for i in range(n):
            a, b = int(ido[i]), int(ido[i + 1])
            out = ctypes.c_void_p()
            ln = ctypes.c_size_t()
            rc = L.oracle_decode_ids(self.h, ids.ctypes.data + 4 * a, b - a, ctypes.byref(out), ctypes.byref(ln))
            if rc:
                raise RuntimeError(f"oracle_decode_ids failed ({rc}) at list {i}")
            parts.append(ctypes.string_at(out, ln.value))
            L.oracle_free(out)
            to[i + 1] = to[i] + np.uint64(ln.value)
        return np.frombuffer(b"".join(parts), dtype=np.uint8), to


class RefModel:
    """The unmodified reference through oracle/ref_shim.cc."""

    def __init__(self, model_bytes):
        if not ref_available():
            raise RuntimeError("oracle/_ref is not built (run `make -C oracle ref` in the dev container)")
        self.lib = L = ctypes.CDLL(_REF_SO)
        L.ref_load_serialized.restype = ctypes.c_void_p
        L.ref_load_serialized.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.ref_free.argtypes = [ctypes.c_void_p]
        L.ref_free_buf.argtypes = [ctypes.c_void_p]
        L.ref_set_vocabulary.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        L.ref_encode_ids.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]
        L.ref_encode_count.restype = ctypes.c_uint64
        L.ref_encode_count.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        L.ref_normalize.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p),
                                    ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_void_p),
                                    ctypes.POINTER(ctypes.c_size_t)]
        L.ref_encode_pieces.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p),
                                        ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
        self.h = L.ref_load_serialized(model_bytes, len(model_bytes))
        if not self.h:
            raise ValueError("reference failed to load the model")

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_free(self.h)
            self.h = None

    def set_vocabulary(self, valid):
        v = [x.encode() if isinstance(x, str) else x for x in valid]
        rc = self.lib.ref_set_vocabulary(self.h, b"\0".join(v) + b"\0", len(v))
        if rc:
            raise RuntimeError("SetVocabulary failed")

    def reset_vocabulary(self):
        self.lib.ref_set_vocabulary(self.h, b"", 0)

    def encode_batch(self, buf, offs, threads=1):
        n = len(offs) - 1
        ido = np.zeros(n + 1, dtype=np.uint64)
        ids = ctypes.c_void_p()
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        rc = self.lib.ref_encode_ids(self.h, buf.ctypes.data, offs.ctypes.data, n, threads, ctypes.byref(ids),
                                     ido.ctypes.data)
        if rc:
            raise RuntimeError(f"reference Encode failed at sentence {rc - 1}")
        a = _i32(ids, int(ido[n]))
        self.lib.ref_free_buf(ids)
        return a, ido

    def encode(self, s):
        buf, offs = pack([s])
        ids, _ = self.encode_batch(buf, offs)
        return ids

    def decode_batch(self, ids, ido, threads=1):
        L = self.lib
        L.ref_decode_ids.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        ido = np.ascontiguousarray(ido, dtype=np.uint64)
        n = len(ido) - 1
        to = np.zeros(n + 1, dtype=np.uint64)
        out = ctypes.c_void_p()
        rc = L.ref_decode_ids(self.h, ids.ctypes.data, ido.ctypes.data, n, threads, ctypes.byref(out), to.ctypes.data)
        if rc:
            raise RuntimeError(f"reference Decode failed at list {rc - 1}")
        text = np.frombuffer(ctypes.string_at(out, int(to[n])), dtype=np.uint8)
        L.ref_free_buf(out)
        return text, to

    def encode_count(self, buf, offs, threads):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        return int(self.lib.ref_encode_count(self.h, buf.ctypes.data, offs.ctypes.data, len(offs) - 1, threads))

    def normalize(self, s):
        out = ctypes.c_void_p()
        n = ctypes.c_size_t()
        n2o = ctypes.c_void_p()
        n2on = ctypes.c_size_t()
        rc = self.lib.ref_normalize(self.h, s, len(s), ctypes.byref(out), ctypes.byref(n), ctypes.byref(n2o),
                                    ctypes.byref(n2on))
        if rc:
            raise RuntimeError("reference Normalize failed")
        res = ctypes.string_at(out, n.value) if n.value else b""
        m = [int(x) for x in _u64(n2o, n2on.value)]
        self.lib.ref_free_buf(out)
        self.lib.ref_free_buf(n2o)
        return res, m

    def nbest_encode(self, s, nbest_size):
        L = self.lib
        L.ref_nbest_encode.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                       ctypes.POINTER(ctypes.c_void_p)]
        ids, co, sc = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        k = L.ref_nbest_encode(self.h, s, len(s), nbest_size, ctypes.byref(ids), ctypes.byref(co), ctypes.byref(sc))
        if k < 0:
            raise RuntimeError("reference NBestEncode failed")
        off = _u32(co, k + 1)
        allids = _i32(ids, int(off[k]))
        scores = np.ctypeslib.as_array(ctypes.cast(sc, ctypes.POINTER(ctypes.c_float)), (max(k, 1),))[:k].copy()
        for ptr in (ids, co, sc):
            L.ref_free_buf(ptr)
        return [allids[int(off[i]):int(off[i + 1])] for i in range(k)], scores

    def sample_encode_batch(self, buf, offs, nbest_size, alpha, seed):
        L = self.lib
        L.ref_sample_encode_ids.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                            ctypes.c_int, ctypes.c_float, ctypes.c_uint, ctypes.POINTER(ctypes.c_void_p),
                                            ctypes.c_void_p]
        n = len(offs) - 1
        ido = np.zeros(n + 1, dtype=np.uint64)
        ids = ctypes.c_void_p()
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        rc = L.ref_sample_encode_ids(self.h, buf.ctypes.data, offs.ctypes.data, n, nbest_size, alpha, seed,
                                     ctypes.byref(ids), ido.ctypes.data)
        if rc:
            raise RuntimeError(f"reference SampleEncode failed at sentence {rc - 1}")
        a = _i32(ids, int(ido[n]))
        L.ref_free_buf(ids)
        return a, ido

    def entropy_batch(self, buf, offs, alpha):
        L = self.lib
        L.ref_entropy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float,
                                  ctypes.c_void_p]
        n = len(offs) - 1
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        ent = np.zeros(n, dtype=np.float32)
        rc = L.ref_entropy(self.h, buf.ctypes.data, offs.ctypes.data, n, alpha, ent.ctypes.data)
        if rc:
            raise RuntimeError(f"reference CalculateEntropy failed at sentence {rc - 1}")
        return ent

    def sample_score_batch(self, buf, offs, samples, alpha, seed, wor=False, include_best=False):
        L = self.lib
        L.ref_sample_score.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                       ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_uint,
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_void_p]
        n = len(offs) - 1
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        co = np.zeros(n * samples + 1, dtype=np.uint64)
        sc = np.zeros(n * samples, dtype=np.float32)
        ids = ctypes.c_void_p()
        rc = L.ref_sample_score(self.h, buf.ctypes.data, offs.ctypes.data, n, samples, alpha, int(wor), int(include_best),
                                seed, ctypes.byref(ids), co.ctypes.data, sc.ctypes.data)
        if rc:
            raise RuntimeError("reference SampleEncodeAndScore failed")
        a = _i32(ids, int(co[-1]))
        self.lib.ref_free_buf(ids)
        return a, co, sc

    def encode_pieces(self, s):
        out = ctypes.c_void_p()
        n = ctypes.c_size_t()
        k = ctypes.c_size_t()
        rc = self.lib.ref_encode_pieces(self.h, s, len(s), ctypes.byref(out), ctypes.byref(n), ctypes.byref(k))
        if rc:
            raise RuntimeError("reference EncodeAsPieces failed")
        raw = ctypes.string_at(out, n.value) if n.value else b""
        self.lib.ref_free_buf(out)
        return raw.split(b"\0")[:k.value]

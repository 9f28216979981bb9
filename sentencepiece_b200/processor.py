"""Host-side mirror of the reference interface for the batched-encode path."""
import ctypes

import numpy as np

from . import _capi


def pack_sentences(sentences):
    """list[bytes|str] -> (uint8[total], uint64[n+1])"""
    bs = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in sentences]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, np.uint8)
    return buf, offs


class Engine:
    """One engine = one model on one GPU (spm_engine in include/spm_b200.h)."""

    def __init__(self, model_bytes, device=0):
        self._lib = _capi.load()
        h = ctypes.c_void_p()
        rc = self._lib.spm_engine_create_from_serialized(model_bytes, len(model_bytes), device, ctypes.byref(h))
        if rc:
            raise RuntimeError(f"spm_engine_create failed ({rc}): {self._lib.spm_last_error(None).decode()}")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.spm_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise RuntimeError(f"spm_b200 error {rc}: {self._lib.spm_last_error(self._h).decode()}")

    def info(self):
        i = _capi.EngineInfo()
        self._check(self._lib.spm_engine_get_info(self._h, ctypes.byref(i)))
        return i

    def set_tuning(self, lanes=0, cap=0, ctas=0):
        self._check(self._lib.spm_engine_set_tuning(self._h, lanes, cap, ctas))

    def set_types(self, types):
        t = np.ascontiguousarray(types, dtype=np.uint8)
        self._check(self._lib.spm_engine_set_types(self._h, t.ctypes.data))

    def cache_reset(self):
        """empties the engine's memo tables (the BPE word cache); results never depend on them"""
        self._check(self._lib.spm_engine_cache_reset(self._h))

    def encode_packed(self, buf, offs, copy=True):
        """Batch encode of a packed host buffer -> (ids int32[], id_offsets uint64[n+1])."""
        n = len(offs) - 1
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        ids = ctypes.c_void_p()
        ido = ctypes.c_void_p()
        self._check(self._lib.spm_encode_ids(self._h, buf.ctypes.data, offs.ctypes.data, n, ctypes.byref(ids),
                                             ctypes.byref(ido)))
        o = np.ctypeslib.as_array(ctypes.cast(ido, ctypes.POINTER(ctypes.c_uint64)), (n + 1,))
        tot = int(o[n])
        a = np.ctypeslib.as_array(ctypes.cast(ids, ctypes.POINTER(ctypes.c_int32)), (max(tot, 1),))[:tot]
        return (a.copy(), o.copy()) if copy else (a, o)

    def encode_packed_ptr(self, bytes_ptr, offs_ptr, n):
        """Raw-pointer variant for benchmarks (pinned host memory): returns total ids."""
        ids = ctypes.c_void_p()
        ido = ctypes.c_void_p()
        self._check(self._lib.spm_encode_ids(self._h, bytes_ptr, offs_ptr, n, ctypes.byref(ids), ctypes.byref(ido)))
        o = np.ctypeslib.as_array(ctypes.cast(ido, ctypes.POINTER(ctypes.c_uint64)), (n + 1,))
        return int(o[n]), ids.value, ido.value

    def encode_spans(self, buf, offs):
        """-> dict(ids, tok_end, id_offsets, normalized(bytes), norm_offsets, n2o)"""
        n = len(offs) - 1
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        p = [ctypes.c_void_p() for _ in range(6)]
        self._check(self._lib.spm_encode_spans(self._h, buf.ctypes.data, offs.ctypes.data, n, *[ctypes.byref(x) for x in p]))
        ido = np.ctypeslib.as_array(ctypes.cast(p[2], ctypes.POINTER(ctypes.c_uint64)), (n + 1,)).copy()
        tot = int(ido[n])
        no = np.ctypeslib.as_array(ctypes.cast(p[4], ctypes.POINTER(ctypes.c_uint64)), (n + 1,)).copy()
        tn = int(no[n])
        ids = np.ctypeslib.as_array(ctypes.cast(p[0], ctypes.POINTER(ctypes.c_int32)), (max(tot, 1),))[:tot].copy()
        te = np.ctypeslib.as_array(ctypes.cast(p[1], ctypes.POINTER(ctypes.c_uint32)), (max(tot, 1),))[:tot].copy()
        norm = ctypes.string_at(p[3], tn) if tn else b""
        n2o = np.ctypeslib.as_array(ctypes.cast(p[5], ctypes.POINTER(ctypes.c_uint32)), (tn + n,)).copy()
        return dict(ids=ids, tok_end=te, id_offsets=ido, normalized=norm, norm_offsets=no, n2o=n2o)

    def decode_packed(self, ids, id_offsets):
        """Batch Decode of packed id lists -> (text uint8[], text_offsets uint64[n+1])."""
        n = len(id_offsets) - 1
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        ido = np.ascontiguousarray(id_offsets, dtype=np.uint64)
        text = ctypes.c_void_p()
        to = ctypes.c_void_p()
        self._check(self._lib.spm_decode_ids(self._h, ids.ctypes.data, ido.ctypes.data, n, ctypes.byref(text), ctypes.byref(to)))
        o = np.ctypeslib.as_array(ctypes.cast(to, ctypes.POINTER(ctypes.c_uint64)), (n + 1,)).copy()
        tot = int(o[n])
        t = np.frombuffer(ctypes.string_at(text, tot), dtype=np.uint8) if tot else np.zeros(0, np.uint8)
        return t, o

    def set_random_seed(self, seed):
        self._check(self._lib.spm_set_random_seed(self._h, seed))

    def nbest_encode(self, buf, offs, nbest_size):
        """-> dict(ids, cand_offsets uint64[n*K+1], scores float32[n*K], n_cands uint32[n], K)"""
        n = len(offs) - 1
        K = max(1, min(nbest_size, 1024))
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        p = [ctypes.c_void_p() for _ in range(4)]
        self._check(self._lib.spm_nbest_encode(self._h, buf.ctypes.data, offs.ctypes.data, n, nbest_size,
                                               *[ctypes.byref(x) for x in p]))
        nc = n * K
        co = np.ctypeslib.as_array(ctypes.cast(p[1], ctypes.POINTER(ctypes.c_uint64)), (nc + 1,)).copy()
        tot = int(co[nc])
        ids = np.ctypeslib.as_array(ctypes.cast(p[0], ctypes.POINTER(ctypes.c_int32)), (max(tot, 1),))[:tot].copy()
        sc = np.ctypeslib.as_array(ctypes.cast(p[2], ctypes.POINTER(ctypes.c_float)), (max(nc, 1),))[:nc].copy()
        nk = np.ctypeslib.as_array(ctypes.cast(p[3], ctypes.POINTER(ctypes.c_uint32)), (max(n, 1),))[:n].copy()
        return dict(ids=ids, cand_offsets=co, scores=sc, n_cands=nk, K=K)

    def sample_encode(self, buf, offs, nbest_size, alpha):
        """SampleEncode over a packed batch -> (ids, id_offsets)"""
        n = len(offs) - 1
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        ids = ctypes.c_void_p()
        ido = ctypes.c_void_p()
        self._check(self._lib.spm_sample_encode_ids(self._h, buf.ctypes.data, offs.ctypes.data, n, nbest_size, alpha,
                                                    ctypes.byref(ids), ctypes.byref(ido)))
        o = np.ctypeslib.as_array(ctypes.cast(ido, ctypes.POINTER(ctypes.c_uint64)), (n + 1,)).copy()
        tot = int(o[n])
        a = np.ctypeslib.as_array(ctypes.cast(ids, ctypes.POINTER(ctypes.c_int32)), (max(tot, 1),))[:tot].copy()
        return a, o

    def calculate_entropy(self, buf, offs, alpha):
        """CalculateEntropy per sentence of a packed batch -> float32[n]"""
        n = len(offs) - 1
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        ent = ctypes.c_void_p()
        self._check(self._lib.spm_calculate_entropy(self._h, buf.ctypes.data, offs.ctypes.data, n, alpha, ctypes.byref(ent)))
        return np.ctypeslib.as_array(ctypes.cast(ent, ctypes.POINTER(ctypes.c_float)), (max(n, 1),))[:n].copy()

    def sample_encode_and_score(self, buf, offs, num_samples, alpha, wor=False, include_best=False):
        """-> (ids, cand_offsets uint64[n*num_samples+1], scores float32[n*num_samples])"""
        n = len(offs) - 1
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        p = [ctypes.c_void_p() for _ in range(3)]
        self._check(self._lib.spm_sample_encode_and_score(self._h, buf.ctypes.data, offs.ctypes.data, n, num_samples, alpha,
                                                          int(wor), int(include_best), *[ctypes.byref(x) for x in p]))
        nc = n * num_samples
        co = np.ctypeslib.as_array(ctypes.cast(p[1], ctypes.POINTER(ctypes.c_uint64)), (nc + 1,)).copy()
        tot = int(co[nc])
        ids = np.ctypeslib.as_array(ctypes.cast(p[0], ctypes.POINTER(ctypes.c_int32)), (max(tot, 1),))[:tot].copy()
        sc = np.ctypeslib.as_array(ctypes.cast(p[2], ctypes.POINTER(ctypes.c_float)), (max(nc, 1),))[:nc].copy()
        return ids, co, sc

    def encode_device(self, d_bytes_ptr, d_offs_ptr, n, total_bytes, d_ids_ptr, ids_cap, d_id_offs_ptr, stream=None):
        """Device-resident batch (pointers are CUDA device pointers, e.g. torch data_ptr())."""
        tot = ctypes.c_uint64()
        self._check(self._lib.spm_encode_ids_device(self._h, d_bytes_ptr, d_offs_ptr, n, total_bytes, d_ids_ptr,
                                                    ids_cap, d_id_offs_ptr, ctypes.byref(tot), stream))
        return tot.value


class SentencePieceProcessor:
    """Mirror of the reference's encode API (same method names / argument meaning as
    python/src/sentencepiece/__init__.py and src/sentencepiece_processor.h for this path)."""

    def __init__(self, model_file=None, model_proto=None, device=0):
        self._engine = None
        if model_file is not None:
            self.Load(model_file, device=device)
        elif model_proto is not None:
            self.LoadFromSerializedProto(model_proto, device=device)

    # sentencepiece_processor.h:245
    def Load(self, model_file, device=0):
        with open(model_file, "rb") as f:
            return self.LoadFromSerializedProto(f.read(), device=device)

    # sentencepiece_processor.h:261
    def LoadFromSerializedProto(self, serialized, device=0):
        self._model_bytes = bytes(serialized)
        self._engine = Engine(self._model_bytes, device=device)
        self._pieces = None
        self._unk_id = -1
        return True

    def _require(self):
        if self._engine is None:
            raise RuntimeError("Model is not initialized.")  # sentencepiece_processor.cc:293-299

    # sentencepiece_processor.h:458 + python batch entry sentencepiece.i:439-446
    def DecodeIds(self, input):
        """Decode(ids) (src/sentencepiece_processor.h, python __init__.py DecodeIds): one id list or a list of lists."""
        self._require()
        single = len(input) == 0 or not isinstance(input[0], (list, tuple, np.ndarray))
        lists = [input] if single else list(input)
        ido = np.zeros(len(lists) + 1, dtype=np.uint64)
        if lists:
            ido[1:] = np.cumsum([len(x) for x in lists], dtype=np.uint64)
        ids = np.concatenate([np.asarray(x, dtype=np.int32) for x in lists]) if int(ido[-1]) else np.zeros(0, np.int32)
        text, to = self._engine.decode_packed(ids, ido)
        raw = text.tobytes()
        out = [raw[int(to[i]):int(to[i + 1])].decode("utf-8", errors="replace") for i in range(len(lists))]
        return out[0] if single else out

    def EncodeAsIds(self, input):
        self._require()
        single = isinstance(input, (str, bytes))
        buf, offs = pack_sentences([input] if single else input)
        ids, ido = self._engine.encode_packed(buf, offs)
        out = [ids[int(ido[i]):int(ido[i + 1])].tolist() for i in range(len(offs) - 1)]
        return out[0] if single else out

    # ---- pieces: ids + token ends in the normalized text (spm_encode_spans); unknown tokens keep their surface
    #      (sentencepiece_processor.cc:609-621) ----
    def EncodeAsPieces(self, input):
        self._require()
        single = isinstance(input, (str, bytes))
        buf, offs = pack_sentences([input] if single else input)
        r = self._engine.encode_spans(buf, offs)
        if self._pieces is None:
            from . import _modelinfo
            self._pieces, self._unk_id = _modelinfo.pieces_and_unk(self._model_bytes)
        ids, te, ido, no, norm = r["ids"], r["tok_end"], r["id_offsets"], r["norm_offsets"], r["normalized"]
        out = []
        for i in range(len(offs) - 1):
            base, begin, row = int(no[i]), 0, []
            for k in range(int(ido[i]), int(ido[i + 1])):
                end = int(te[k])
                if int(ids[k]) == self._unk_id:
                    row.append(norm[base + begin:base + end].decode("utf-8", errors="replace"))
                else:
                    row.append(self._pieces[int(ids[k])])
                begin = end
            out.append(row)
        return out[0] if single else out

    def encode(self, input, out_type=int, enable_sampling=False, nbest_size=-1, alpha=0.1):
        """python/src/sentencepiece/__init__.py Encode: out_type int | str; enable_sampling uses SampleEncode."""
        if enable_sampling:
            if out_type is not int:
                raise NotImplementedError("sampled pieces: use the C++ host layer (SampleEncodeAsPieces)")
            return self.SampleEncodeAsIds(input, nbest_size, alpha)
        return self.EncodeAsIds(input) if out_type is int else self.EncodeAsPieces(input)

    Encode = encode

    def _split(self, input):
        single = isinstance(input, (str, bytes))
        return single, pack_sentences([input] if single else input)

    # sentencepiece_processor.h:481; the draws of a batch are taken in order on one generator (SetRandomGeneratorSeed)
    def SampleEncodeAsIds(self, input, nbest_size, alpha):
        self._require()
        single, (buf, offs) = self._split(input)
        ids, ido = self._engine.sample_encode(buf, offs, nbest_size, alpha)
        out = [ids[int(ido[i]):int(ido[i + 1])].tolist() for i in range(len(offs) - 1)]
        return out[0] if single else out

    # sentencepiece_processor.h:471
    def NBestEncodeAsIds(self, input, nbest_size):
        self._require()
        single, (buf, offs) = self._split(input)
        r = self._engine.nbest_encode(buf, offs, nbest_size)
        K, co = r["K"], r["cand_offsets"]
        out = [[r["ids"][int(co[i * K + c]):int(co[i * K + c + 1])].tolist() for c in range(int(r["n_cands"][i]))]
               for i in range(len(offs) - 1)]
        return out[0] if single else out

    # sentencepiece_processor.h:521-526
    def CalculateEntropy(self, input, alpha):
        self._require()
        single, (buf, offs) = self._split(input)
        ent = self._engine.calculate_entropy(buf, offs, alpha).tolist()
        return ent[0] if single else ent

    def SetRandomGeneratorSeed(self, seed):
        self._require()
        self._engine.set_random_seed(seed)

    @property
    def engine(self):
        self._require()
        return self._engine

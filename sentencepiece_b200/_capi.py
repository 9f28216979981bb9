"""ctypes binding of include/spm_b200.h (the C-ABI shared library built in-tree).

There is deliberately no fallback: if libspm_b200.so is missing or no B200 is
visible, loading / engine creation raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libspm_b200.so")

# every symbol include/spm_b200.h declares
EXPORTS = [
    "spm_engine_create", "spm_engine_create_from_serialized", "spm_engine_destroy", "spm_engine_set_types",
    "spm_last_error", "spm_encode_ids", "spm_encode_spans", "spm_encode_ids_device", "spm_host_alloc",
    "spm_host_free", "spm_engine_get_info", "spm_engine_set_tuning", "spm_nbest_encode", "spm_set_random_seed",
    "spm_sample_encode_ids", "spm_decode_ids", "spm_engine_set_unk_surface", "spm_calculate_entropy",
    "spm_sample_encode_and_score", "spm_engine_cache_reset",
]


class ModelDesc(ctypes.Structure):
    _fields_ = [("model_type", ctypes.c_int32), ("vocab_size", ctypes.c_int32),
                ("piece_bytes", ctypes.c_char_p), ("piece_off", ctypes.c_void_p),
                ("scores", ctypes.c_void_p), ("types", ctypes.c_void_p),
                ("byte_fallback", ctypes.c_uint8), ("treat_whitespace_as_suffix", ctypes.c_uint8),
                ("add_dummy_prefix", ctypes.c_uint8), ("remove_extra_whitespaces", ctypes.c_uint8),
                ("escape_whitespaces", ctypes.c_uint8), ("reserved_", ctypes.c_uint8 * 3),
                ("charsmap", ctypes.c_char_p), ("charsmap_bytes", ctypes.c_size_t)]


class EngineInfo(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("sm_count", ctypes.c_int32), ("model_type", ctypes.c_int32),
                ("vocab_size", ctypes.c_int32), ("unk_id", ctypes.c_int32), ("min_score", ctypes.c_float),
                ("max_score", ctypes.c_float), ("trie_units", ctypes.c_uint32), ("trie_hot_units", ctypes.c_uint32),
                ("charsmap_units", ctypes.c_uint32), ("last_kernel_launches", ctypes.c_uint64),
                ("last_kernel_ms", ctypes.c_float), ("last_main_kernel_ms", ctypes.c_float),
                ("last_h2d_bytes", ctypes.c_uint64), ("last_d2h_bytes", ctypes.c_uint64),
                ("last_deferred", ctypes.c_uint64)]


_lib = None


def load():
    """Loads libspm_b200.so (raises if it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: the CUDA engine has not been built "
                           "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, cp, sz = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t
    P = ctypes.POINTER
    L.spm_engine_create.argtypes = [P(ModelDesc), ctypes.c_int, P(vp)]
    L.spm_engine_create_from_serialized.argtypes = [cp, sz, ctypes.c_int, P(vp)]
    L.spm_engine_destroy.argtypes = [vp]
    L.spm_engine_destroy.restype = None
    L.spm_engine_set_types.argtypes = [vp, vp]
    L.spm_engine_cache_reset.argtypes = [vp]
    L.spm_last_error.argtypes = [vp]
    L.spm_last_error.restype = cp
    L.spm_encode_ids.argtypes = [vp, vp, vp, sz, P(vp), P(vp)]
    L.spm_encode_spans.argtypes = [vp, vp, vp, sz, P(vp), P(vp), P(vp), P(vp), P(vp), P(vp)]
    L.spm_encode_ids_device.argtypes = [vp, vp, vp, sz, ctypes.c_uint64, vp, ctypes.c_uint64, vp,
                                        P(ctypes.c_uint64), vp]
    L.spm_host_alloc.argtypes = [sz]
    L.spm_host_alloc.restype = vp
    L.spm_host_free.argtypes = [vp]
    L.spm_host_free.restype = None
    L.spm_engine_get_info.argtypes = [vp, P(EngineInfo)]
    L.spm_engine_set_tuning.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.spm_nbest_encode.argtypes = [vp, vp, vp, sz, ctypes.c_int, P(vp), P(vp), P(vp), P(vp)]
    L.spm_set_random_seed.argtypes = [vp, ctypes.c_uint32]
    L.spm_sample_encode_ids.argtypes = [vp, vp, vp, sz, ctypes.c_int, ctypes.c_float, P(vp), P(vp)]
    L.spm_decode_ids.argtypes = [vp, vp, vp, sz, P(vp), P(vp)]
    L.spm_engine_set_unk_surface.argtypes = [vp, cp, sz]
    L.spm_calculate_entropy.argtypes = [vp, vp, vp, sz, ctypes.c_float, P(vp)]
    L.spm_sample_encode_and_score.argtypes = [vp, vp, vp, sz, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                              P(vp), P(vp), P(vp)]
    _lib = L
    return L

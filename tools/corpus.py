"""ctypes wrapper around tools/corpus_gen.c (seeded synthetic corpora).

Data plumbing shared by bench.py, the parity tests and tools/make_models.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "libcorpus_gen.so")
KINDS = {"en": 0, "mixed": 1}


def build(force=False):
    src = os.path.join(_HERE, "corpus_gen.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


class CorpusGen:
    def __init__(self):
        self.lib = ctypes.CDLL(build())
        L = self.lib
        L.corpus_gen_create.restype = ctypes.c_void_p
        L.corpus_gen_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.corpus_gen_destroy.argtypes = [ctypes.c_void_p]
        L.corpus_gen_max_sentence_bytes.restype = ctypes.c_uint64
        L.corpus_gen_max_sentence_bytes.argtypes = [ctypes.c_int]
        L.corpus_gen_fill.restype = ctypes.c_uint64
        L.corpus_gen_fill.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64,
                                      ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
        gold = os.path.join(_ROOT, "tests", "golden")
        self.h = L.corpus_gen_create(os.path.join(gold, "en_wordlist.tsv").encode(),
                                     os.path.join(gold, "ja_charlist.tsv").encode())
        if not self.h:
            raise RuntimeError("corpus_gen_create failed (word lists missing?)")

    def fill(self, kind, seed, n, first=0, out=None):
        """Returns (bytes uint8[total], offsets uint64[n+1]) for sentences [first, first+n).

        `out` may be a preallocated uint8 numpy array (e.g. a view of pinned memory)."""
        k = KINDS[kind] if isinstance(kind, str) else int(kind)
        cap = int(self.lib.corpus_gen_max_sentence_bytes(k)) * (n + 1)
        offs = np.empty(n + 1, dtype=np.uint64)
        if out is None:
            buf = np.empty(cap, dtype=np.uint8)
        else:
            buf = out
            cap = buf.size
        total = self.lib.corpus_gen_fill(self.h, k, seed, first, n, buf.ctypes.data, cap, offs.ctypes.data)
        if total == 2**64 - 1:
            raise RuntimeError("corpus buffer too small")
        return (buf[:total] if out is None else buf[:total]), offs

    def lines(self, kind, seed, n, first=0):
        b, o = self.fill(kind, seed, n, first)
        raw = b.tobytes()
        return [raw[int(o[i]):int(o[i + 1])] for i in range(n)]

    def write_file(self, path, kind, seed, n, first=0):
        b, o = self.fill(kind, seed, n, first)
        raw = b.tobytes()
        with open(path, "wb") as f:
            for i in range(n):
                f.write(raw[int(o[i]):int(o[i + 1])])
                f.write(b"\n")

    def __del__(self):
        try:
            self.lib.corpus_gen_destroy(self.h)
        except Exception:
            pass

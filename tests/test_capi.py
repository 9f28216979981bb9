"""The C-ABI shared library: it loads, exports every symbol include/spm_b200.h declares,
and -- on a box without a GPU -- refuses to create an engine instead of falling back.  CPU only."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, model_bytes
from sentencepiece_b200 import _capi


def test_header_symbols_are_exported():
    lib = _capi.load()
    hdr = open(os.path.join(ROOT, "include", "spm_b200.h")).read()
    declared = set(re.findall(r"\b(spm_[a-z_]+)\s*\(", hdr))
    declared -= {"spm_engine", "spm_model_desc", "spm_engine_info"}
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here; the failure path is for GPU-less boxes")
    lib = _capi.load()
    h = ctypes.c_void_p()
    mb = model_bytes("botchan8k")
    rc = lib.spm_engine_create_from_serialized(mb, len(mb), 0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.spm_last_error(None) or b"CUDA" in lib.spm_last_error(None)


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "sentencepiece_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="replace").read()
                assert "oracle/" not in txt.replace("see oracle/Makefile", "") and "import oracle" not in txt \
                    and "from oracle" not in txt, os.path.join(d, f)


def test_makefile_tracks_every_kernel_header():
    """The engine is one translation unit that includes every kernel header: the Makefile rule must depend on all of
    them (a missing header dependency once made `make` a silent no-op for three kernel experiments)."""
    csrc = os.path.join(ROOT, "sentencepiece_b200", "csrc")
    mk = open(os.path.join(csrc, "Makefile")).read()
    rule = re.search(r"^\$\(OUT\)/engine\.o:(.*)$", mk, re.M).group(1)
    assert "$(wildcard *.cuh)" in rule and "$(wildcard *.h)" in rule

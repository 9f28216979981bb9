"""Host-side logic of the Python mirror that needs no GPU: the piece-string reader (_modelinfo.py) against the
reference's own Python binding (the `sentencepiece` wheel of this image, VERSION 0.2.1 == /root/reference's), and the
mirror's behaviour without a device: it must fail loudly, never fall back.  CPU only."""
import os

import numpy as np
import pytest

from conftest import ROOT, model_bytes
from sentencepiece_b200 import _modelinfo

MODELS = ["uni32k", "bpe32k", "mix_bf8k", "mix_bpe4k", "botchan8k"]


@pytest.mark.parametrize("model", MODELS)
def test_piece_reader_matches_the_reference_binding(model):
    spm = pytest.importorskip("sentencepiece")
    mb = model_bytes(model)
    sp = spm.SentencePieceProcessor(model_proto=mb)
    pieces, unk = _modelinfo.pieces_and_unk(mb)
    assert len(pieces) == sp.get_piece_size()
    assert unk == sp.unk_id()
    assert pieces == [sp.id_to_piece(i) for i in range(sp.get_piece_size())]


def test_piece_reader_rejects_garbage():
    with pytest.raises((ValueError, IndexError)):
        _modelinfo.pieces_and_unk(b"\x0b\xff\xff\xff")  # wire type 3 (group) is not part of a ModelProto


def test_mirror_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here; the failure path is for GPU-less boxes")
    from sentencepiece_b200 import SentencePieceProcessor
    with pytest.raises(RuntimeError, match="CUDA|fallback"):
        SentencePieceProcessor(model_file=os.path.join(ROOT, "tests", "golden", "models", "botchan8k.model"))
    sp = SentencePieceProcessor()          # not loaded: the reference's "Model is not initialized." (sentencepiece_processor.cc:293-299)
    with pytest.raises(RuntimeError, match="not initialized"):
        sp.DecodeIds([1, 2, 3])
    with pytest.raises(RuntimeError, match="not initialized"):
        sp.EncodeAsIds("hello")


def test_shard_ranges_cover_and_balance():
    """sharding.shard_ranges: contiguous, complete, byte-balanced ranges"""
    from sentencepiece_b200.sharding import shard_ranges
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 400, size=10007).astype(np.uint64)
    offs = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    for world in (1, 2, 3, 8):
        r = shard_ranges(offs, world)
        assert len(r) == world and r[0][0] == 0 and r[-1][1] == len(lens)
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        sizes = [int(offs[hi] - offs[lo]) for lo, hi in r]
        assert max(sizes) - min(sizes) <= 2 * 400 + int(offs[-1]) // (50 * world)

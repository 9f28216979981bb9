"""The Python mirror (sentencepiece_b200.SentencePieceProcessor) against the reference's own Python layer (the
`sentencepiece` wheel of this image, version == the reference's VERSION.txt): EncodeAsIds / EncodeAsPieces / encode(out_type),
DecodeIds, NBestEncodeAsIds, CalculateEntropy, and sampling determinism.  Needs a B200."""
import os

import numpy as np
import pytest

from conftest import MODELS_DIR

pytestmark = pytest.mark.gpu
spm = pytest.importorskip("sentencepiece")


@pytest.mark.parametrize("model,kind", [("uni32k", "en"), ("mix_bf8k", "mixed"), ("bpe32k", "en"), ("botchan8k", "mixed")])
def test_python_mirror_matches_reference_wheel(model, kind, corpus_gen):
    from sentencepiece_b200 import SentencePieceProcessor
    path = os.path.join(MODELS_DIR, model + ".model")
    ours = SentencePieceProcessor(model_file=path)
    ref = spm.SentencePieceProcessor(model_file=path)
    lines = [l.decode("utf-8", "replace") for l in corpus_gen.lines(kind, 4401, 2000)]
    lines += ["", "   ", "hello world", "\U0001F600 unknown あい", "x"]
    assert ours.EncodeAsIds(lines) == ref.encode(lines, out_type=int)
    assert ours.EncodeAsPieces(lines) == ref.encode(lines, out_type=str)
    assert ours.encode(lines[7], out_type=str) == ref.encode(lines[7], out_type=str)
    assert ours.encode(lines[7]) == ref.encode(lines[7])
    ids = ref.encode(lines, out_type=int)
    assert ours.DecodeIds(ids) == ref.decode(ids)
    assert ours.DecodeIds(ids[3]) == ref.decode(ids[3])
    if model != "bpe32k":
        few = lines[:200]
        assert ours.NBestEncodeAsIds(few, 5) == [ref.nbest_encode(s, nbest_size=5, out_type=int) for s in few]
        ent = ours.CalculateEntropy(few, 0.5)
        exp = [ref.calculate_entropy(s, 0.5) for s in few]
        np.testing.assert_allclose(ent, exp, rtol=2e-5, atol=2e-5)
        # seeded sampling is reproducible through the mirror, and nbest_size < 0 samples from the whole lattice
        ours.SetRandomGeneratorSeed(11)
        a = ours.SampleEncodeAsIds(few, -1, 0.5)
        ours.SetRandomGeneratorSeed(11)
        assert ours.SampleEncodeAsIds(few, -1, 0.5) == a
        assert a != ours.EncodeAsIds(few)
        assert ours.encode(few, enable_sampling=True, nbest_size=8, alpha=0.5) is not None

"""sentencepiece_b200 -- B200-native batched subword-encode engine.

Python host-side mirror of the reference's encode API for this path
(reference: python/src/sentencepiece/__init__.py:471-560 `Encode`, C++
src/sentencepiece_processor.h:245-460).  Everything here is a thin layer over the
C ABI in include/spm_b200.h; the work happens in hand-written sm_100a kernels
(sentencepiece_b200/csrc).  No CPU fallback exists.
"""
from .processor import SentencePieceProcessor, Engine, pack_sentences  # noqa: F401

__all__ = ["SentencePieceProcessor", "Engine", "pack_sentences"]

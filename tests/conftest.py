import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


MODELS_DIR = os.path.join(ROOT, "tests", "golden", "models")


def model_bytes(name):
    with open(os.path.join(MODELS_DIR, name + ".model"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def corpus_gen():
    import corpus
    return corpus.CorpusGen()

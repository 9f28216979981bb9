"""The drop-in caller: spm_encode_b200 (C++ host layer over the C ABI) must print exactly what the
reference's spm_encode prints for --output_format=id / piece (BASELINE.json config 1 plumbing:
train-on-botchan model, encode a text file, compare by md5).  Needs a B200."""
import hashlib
import os
import subprocess

import pytest

from conftest import ROOT, MODELS_DIR, model_bytes
from oracle import oracle_py

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "sentencepiece_b200", "lib", "spm_encode_b200")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "spm_encode")


def md5(b):
    return hashlib.md5(b).hexdigest()


@pytest.mark.parametrize("model,kind,fmt,extra", [("botchan8k", "en", "id", ""), ("botchan8k", "en", "piece", ""),
                                                  ("mix_bf8k", "mixed", "id", ""), ("mix_bf8k", "mixed", "piece", ""),
                                                  ("uni32k", "mixed", "piece", "bos:eos"), ("uni32k", "mixed", "piece", "unk"),
                                                  ("bpe32k", "en", "id", "reverse:eos"), ("bpe32k", "en", "piece", "")])
def test_cli_matches_reference(model, kind, fmt, extra, corpus_gen, tmp_path):
    assert os.path.exists(CLI), "spm_encode_b200 has not been built (__graft_entry__.build())"
    path = str(tmp_path / "in.txt")
    n = 3000
    # text files cannot carry newlines inside a sentence; the generator never emits them
    corpus_gen.write_file(path, kind, 5151, n)
    mpath = os.path.join(MODELS_DIR, model + ".model")
    args = [f"--model={mpath}", f"--output_format={fmt}", f"--input={path}"]
    if extra:
        args.append(f"--extra_options={extra}")
    ours = subprocess.run([CLI, "--batch_lines=1000"] + args, capture_output=True, check=True).stdout
    if os.path.exists(REF_CLI):
        ref = subprocess.run([REF_CLI] + args, capture_output=True, check=True).stdout
        assert md5(ours) == md5(ref), f"{model} {fmt} {extra}: output differs from the reference spm_encode"
    elif fmt == "id" and not extra:
        om = oracle_py.OracleModel(model_bytes(model))
        lines = [ln for ln in open(path, "rb").read().split(b"\n")][:n]
        exp = b"".join(b" ".join(str(i).encode() for i in om.encode(s)[0]) + b"\n" for s in lines)
        assert md5(ours) == md5(exp)
    else:
        pytest.skip("oracle/_ref/spm_encode not on this box")

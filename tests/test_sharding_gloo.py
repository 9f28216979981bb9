"""N>1 host-side logic on CPU: world_size-2 gloo run of the sharding + id-buffer gather,
with the oracle standing in for the per-rank encoder.  CPU only."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, model_bytes


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import corpus
    from oracle import oracle_py
    from sentencepiece_b200.sharding import gather_ids, shard_ranges
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = corpus.CorpusGen()
    buf, offs = g.fill("mixed", 31337, 600)
    lo, hi = shard_ranges(offs, world)[rank]
    sub_offs = offs[lo:hi + 1]
    om = oracle_py.OracleModel(model_bytes("mix_bf8k"))
    ids, ido = om.encode_batch(buf, sub_offs)  # offsets are absolute into buf: fine for the oracle
    res = gather_ids(torch.from_numpy(ids), torch.from_numpy(ido.astype(np.int64)), dst=0)
    if rank == 0:
        np.savez(out_path, ids=res[0].numpy(), offs=res[1].numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path, corpus_gen):
    from oracle import oracle_py
    from sentencepiece_b200.sharding import shard_ranges
    out = str(tmp_path / "gathered.npz")
    port = 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    buf, offs = corpus_gen.fill("mixed", 31337, 600)
    ids, ido = oracle_py.OracleModel(model_bytes("mix_bf8k")).encode_batch(buf, offs)
    assert np.array_equal(z["ids"], ids) and np.array_equal(z["offs"], ido.astype(np.int64))
    # ranges: contiguous cover, balanced by bytes
    r = shard_ranges(offs, 8)
    assert r[0][0] == 0 and r[-1][1] == 600 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    sizes = [int(offs[h]) - int(offs[l]) for l, h in r]
    assert max(sizes) - min(sizes) < 1024


def test_shard_ranges_edge_cases():
    from sentencepiece_b200.sharding import shard_ranges
    assert shard_ranges(np.array([0], dtype=np.uint64), 4) == [(0, 0)] * 4
    r = shard_ranges(np.array([0, 10], dtype=np.uint64), 4)
    assert r[0][0] == 0 and r[-1][1] == 1 and sum(h - l for l, h in r) == 1
    r = shard_ranges(np.array([5, 5, 5, 5], dtype=np.uint64), 2)  # all-empty sentences
    assert sum(h - l for l, h in r) == 3

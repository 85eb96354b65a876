"""N > 1 host logic on CPU: two gloo ranks shard the pairs, 'stitch' their shard (the oracle stands
in for the engine here — this test is about sharding and the homography all-gather, not kernels)
and all-gather the per-pair records; every rank must end up with every homography."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, n_pairs, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module(PKG + ".dist")
    synth = importlib.import_module(PKG + ".synth")
    from oracle.oracle import Oracle
    O = Oracle()
    mine = d.shard_pairs(n_pairs, rank, world)
    res = []
    for p in mine:
        l, r, _ = synth.make_pair(320, 200, seed=100 + p)
        o = O.stitch_pair(l, r, seed=12345)
        res.append({"H": o["H"], "status": 0 if o["status"] == 1 else 5, "best": o["stats"]["best"]})
    allr = d.all_gather_results(d.pack_results(mine, res), n_pairs)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), allr)
    dist.destroy_process_group()


def test_shard_pairs_round_robin():
    d = importlib.import_module(PKG + ".dist")
    assert d.shard_pairs(10, 0, 4) == [0, 4, 8] and d.shard_pairs(10, 3, 4) == [3, 7]
    allp = sorted(sum((d.shard_pairs(257, r, 8) for r in range(8)), []))
    assert allp == list(range(257))
    assert d.shard_pairs(1, 1, 2) == []


@pytest.mark.parametrize("n_pairs", [5, 2])
def test_two_ranks_gather_all_homographies(tmp_path, n_pairs):
    port = 29000 + os.getpid() % 2000 + n_pairs
    mp.spawn(_worker, args=(2, port, n_pairs, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert a.shape == (n_pairs, 12) and np.array_equal(a, b)
    assert list(a[:, 11]) == list(range(n_pairs))
    # the gathered homographies are the ones a single process computes
    synth = importlib.import_module(PKG + ".synth")
    from oracle.oracle import Oracle
    O = Oracle()
    for p in range(n_pairs):
        l, r, _ = synth.make_pair(320, 200, seed=100 + p)
        o = O.stitch_pair(l, r, seed=12345)
        if o["status"] == 1:
            assert np.array_equal(a[p, :9].view(np.uint64), o["H"].reshape(9).view(np.uint64))

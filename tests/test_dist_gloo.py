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


# ---------------- distributed chain mode on CPU: oracle-backed stand-in for the engine ------------
class _OracleEngine:
    """implements the Engine methods stitch_chain_distributed uses, on the CPU oracle (this test is
    about pair sharding, the record all-gather, identical geometry on all ranks and band tiling)"""

    def __init__(self):
        from oracle.oracle import Oracle
        self.O = Oracle()

    def pairHomography(self, left, right):
        H = self.O.pair_homography(left, right, seed=12345)
        return {"H": H if H is not None else np.zeros((3, 3)), "status": 0 if H is not None else 5, "best": 0}

    def composeChain(self, pair_H):
        Hs = [np.eye(3)]
        for H in pair_H:
            if H is None:
                break
            Hs.append(self.O.mul33(Hs[-1], H))
        return Hs

    def chainGeometry(self, sizes, Hs):
        geom, T = self.O.chain_geometry(sizes, Hs)
        return True, geom, T

    def renderChainBand(self, images, Hs, geom, T, y0, bh, out=None):
        cw, ch, x0, yy0 = geom
        canvas = np.zeros((ch, cw, 3), np.uint8)
        canvas[yy0:yy0 + images[0].shape[0], x0:x0 + images[0].shape[1]] = images[0]
        for im, H in list(zip(images, Hs))[1:]:
            w = self.O.warp_perspective(im, self.O.mul33(T, H), (cw, ch))
            nz = w.any(axis=2)
            canvas[nz] = w[nz]
        if out is not None:
            out[:] = canvas[y0:y0 + bh]
            return out
        return canvas[y0:y0 + bh]


def _chain_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module(PKG + ".dist")
    synth = importlib.import_module(PKG + ".synth")
    views = synth.make_strip(n=4, w=320, h=200, seed=21)
    # the NVLink replication of the inputs (upload 1 / W, all-gather) with CPU tensors: every rank must end up with
    # every image, in order, also when the count is not a multiple of the world size
    for k in (len(views), len(views) - 1):
        rep = d.replicate_images(views[:k], "cpu")
        assert len(rep) == k and all(np.array_equal(t.numpy(), v) for t, v in zip(rep, views[:k]))
    pano, allr = d.stitch_chain_distributed(_OracleEngine(), views)
    np.save(os.path.join(out_dir, "pano%d.npy" % rank), np.array(pano))
    dist.destroy_process_group()


def test_band_rows_cover_canvas():
    d = importlib.import_module(PKG + ".dist")
    for ch, w in ((401, 2), (7, 8), (1000, 3)):
        rows = [d.band_rows(ch, r, w) for r in range(w)]
        assert rows[0][0] == 0 and sum(h for _, h in rows) == ch
        assert all(rows[i][0] + rows[i][1] == rows[i + 1][0] for i in range(w - 1))


def test_distributed_chain_two_ranks(tmp_path):
    port = 31000 + os.getpid() % 2000
    mp.spawn(_chain_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "pano0.npy"), np.load(tmp_path / "pano1.npy")
    assert np.array_equal(a, b)
    from oracle.oracle import Oracle
    synth = importlib.import_module(PKG + ".synth")
    ref, _ = Oracle().stitch_chain(synth.make_strip(n=4, w=320, h=200, seed=21), seed=12345)
    assert a.shape == ref.shape and np.array_equal(a, ref)


# ---- bench.py's chain leg at N > 1: rank 0 runs a child torchrun job, the other ranks wait on the rendezvous store ----
def _bench_mod():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    return bench


_STUB = r'''
import json, os, sys, time
import torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mode = sys.argv[1]
if mode == "hang":
    time.sleep(600)
dist.barrier()
if rank == 0:
    print("noise")
    print(json.dumps({"metric": "stub", "n_gpus": world, "cpus": len(os.sched_getaffinity(0)),
                      "nvlink": os.environ.get("PANO_CHAIN_NVLINK"), "lanes": os.environ.get("PANO_BATCH_LANES"),
                      "pid": os.getpid()}), flush=True)
if mode == "late":
    time.sleep(600)
dist.destroy_process_group()
'''


def _chain_leg_worker(rank, world, port, out_dir, mode):
    sys.path.insert(0, ROOT)
    import json
    import time
    import types
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    # what a torchrun worker of bench.py has in its environment, and a rank that bound itself to one CPU
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), PANO_BATCH_LANES="7", TORCHELASTIC_RUN_ID="x")
    cpus = sorted(os.sched_getaffinity(0))
    os.sched_setaffinity(0, {cpus[rank % len(cpus)]})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bench = _bench_mod()
    stub = os.path.join(out_dir, "stub.py")
    if rank == 0:
        open(stub, "w").write(_STUB)
    dist.barrier()

    def make_cmd(w, c, p):
        cmd = bench.chain_child_command(w, c, p)
        i = cmd.index(os.path.join(ROOT, "bench.py"))
        return cmd[:i] + [stub, mode]
    a = types.SimpleNamespace(extras_timeout={"ok": 60.0, "hang": 8.0, "late": 30.0}[mode])
    t0 = time.time()
    out = bench.chain_config_at_n(a, dist, rank, world, cpus, make_cmd=make_cmd)
    json.dump({"out": out, "seconds": time.time() - t0, "cpus": len(cpus)}, open(os.path.join(out_dir, "r%d_%s.json" % (rank, mode)), "w"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["ok", "hang", "late"])
def test_bench_chain_leg_child_job_and_store_wait(tmp_path, mode):
    import json
    port = 31000 + os.getpid() % 2000 + {"ok": 7, "hang": 11, "late": 13}[mode]
    mp.spawn(_chain_leg_worker, args=(2, port, str(tmp_path), mode), nprocs=2, join=True)
    r0, r1 = [json.load(open(tmp_path / ("r%d_%s.json" % (r, mode)))) for r in (0, 1)]
    assert r1["out"] is None                                   # only rank 0 reports
    if mode == "ok":
        o = r0["out"]
        assert o["metric"] == "stub" and o["n_gpus"] == 2 and "child_seconds" in o
        assert o["cpus"] == r0["cpus"]                         # the affinity before binding is back in the child job
        assert o["nvlink"] == "0" and o["lanes"] is None       # validated upload variant; the parent's lane count is not inherited
        assert r1["seconds"] >= r0["seconds"] - 2.0            # rank 1 waited for rank 0's child job
    elif mode == "late":                                       # measured and printed, stuck on its way out: the line is kept
        assert r0["out"]["metric"] == "stub" and "had not exited after 30 s" in r0["out"]["note"]
        assert r0["seconds"] < 75 and r1["seconds"] < 75
    else:
        assert "killed after 8 s" in r0["out"]["error"]
        assert r0["seconds"] < 40 and r1["seconds"] < 40       # the stuck job is killed as a group; nobody keeps waiting

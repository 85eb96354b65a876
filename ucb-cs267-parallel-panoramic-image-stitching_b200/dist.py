"""Multi-GPU plumbing: one process per GPU, independent pairs sharded over ranks.

The stitching path shards naturally by image pair: pair p goes to rank p mod W,
no data crosses GPUs on the data path.  The only exchange is the all-gather of the small per-pair
results (3x3 homography, status, inlier count: 96 bytes per pair) so that every rank ends up with
every homography — NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import numpy as np

RECORD = 12  # H[0..8], status, inliers, pair index


def shard_pairs(n_pairs, rank, world):
    """indices of the pairs rank `rank` of `world` processes (round robin: p mod W)"""
    return list(range(rank, n_pairs, world))


def pack_results(indices, results):
    """results: list of dicts with 'H' (3x3), 'status', 'best' -> float64 array [n, RECORD]"""
    out = np.zeros((len(indices), RECORD), np.float64)
    for row, (i, r) in enumerate(zip(indices, results)):
        out[row, :9] = np.asarray(r["H"], np.float64).reshape(9)
        out[row, 9] = r["status"]
        out[row, 10] = r["best"]
        out[row, 11] = i
    return out


def all_gather_results(local, n_pairs, device=None, group=None):
    """all-gather of the per-pair records; returns [n_pairs, RECORD] ordered by pair index on
    every rank.  `local` is this rank's pack_results() array.  Works on any torch.distributed
    backend (tensors are moved to `device` for NCCL)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = (n_pairs + world - 1) // world          # shards are padded to the largest one
    buf = torch.full((per, RECORD), -1.0, dtype=torch.float64)
    buf[:len(local)] = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        buf = buf.to(device)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    allr = torch.cat(gathered).cpu().numpy()
    allr = allr[allr[:, 11] >= 0]
    out = np.zeros((n_pairs, RECORD), np.float64)
    out[:, 11] = -1
    for row in allr:
        out[int(row[11])] = row
    assert (out[:, 11] >= 0).all(), "a pair was not reported by any rank"
    return out

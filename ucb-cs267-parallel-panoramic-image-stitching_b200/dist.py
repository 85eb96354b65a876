"""Multi-GPU plumbing: one process per GPU, independent pairs sharded over ranks.

The stitching path shards naturally by image pair: pair p goes to rank p mod W,
no data crosses GPUs on the data path.  The only exchange is the all-gather of the small per-pair
results (3x3 homography, status, inlier count: 96 bytes per pair) so that every rank ends up with
every homography — NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import os

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


def band_rows(canvas_h, rank, world):
    """rows of the canvas rendered by `rank`: contiguous bands, remainder to the first ranks"""
    base, rem = divmod(canvas_h, world)
    y0 = rank * base + min(rank, rem)
    return y0, base + (1 if rank < rem else 0)


class SharedCanvas:
    """A host canvas shared by all ranks of one node ("output bands are written straight to a pinned
    host canvas; no NCCL bulk transfer").  POSIX shared memory under /dev/shm, created by rank 0, mapped by every
    rank and - when a CUDA runtime is around - page-locked with cudaHostRegister so that a band's device -> host
    copy lands in it directly.  Kept and reused across calls while it is large enough."""
    _cache = {}

    def __init__(self, nbytes, group=None):
        import torch.distributed as dist
        rank = dist.get_rank(group)
        cap = max(int(nbytes * 1.25), 1 << 20)
        seq = SharedCanvas._cache.get("seq", 0) + 1
        SharedCanvas._cache["seq"] = seq
        self.path = "/dev/shm/pano_canvas_%s_%d" % (os.environ.get("MASTER_PORT", "0"), seq)
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(cap)
        dist.barrier(group=group)
        self.buf = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(cap,))
        dist.barrier(group=group)
        if rank == 0:
            os.unlink(self.path)          # the mappings keep it alive
        self.cap = cap
        self.registered = False
        try:
            import torch
            if torch.cuda.is_available():
                rc = torch.cuda.cudart().cudaHostRegister(self.buf.ctypes.data, cap, 0)
                self.registered = int(rc) == 0
        except Exception:
            self.registered = False

    @classmethod
    def get(cls, nbytes, group=None):
        c = cls._cache.get(id(group))
        if c is None or c.cap < nbytes:
            c = cls(nbytes, group)
            cls._cache[id(group)] = c
        return c


def replicate_images(images, device, group=None):
    """Hands all `images` (equal shapes, host arrays every rank holds) to every rank on `device` without every rank
    pushing all of them through its own host link: rank r uploads images r, r + W, ... and one
    all-gather (NCCL over NVLink on GPUs) completes the set.  On the 8-GPU boxes of this pool the host side moves
    ~105-130 GB/s in total, so W ranks uploading all images each would spend most of the step there.
    Returns one tensor view per image, in input order."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = len(images)
    per = (n + world - 1) // world
    h, w = np.asarray(images[0]).shape[:2]
    part = torch.zeros((per, h, w, 3), dtype=torch.uint8, device=device)
    for slot, j in enumerate(range(rank, n, world)):
        part[slot].copy_(torch.from_numpy(np.ascontiguousarray(images[j])), non_blocking=True)
    allimg = torch.empty((world * per, h, w, 3), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(allimg, part, group=group)
    if allimg.is_cuda:
        torch.cuda.current_stream().synchronize()     # the engine works on its own stream
    return [allimg[(j % world) * per + j // world] for j in range(n)]


def stitch_chain_distributed(engine, images, device=None, group=None, replicate=None):
    """Chain-mode panorama over all ranks of the process group (SURVEY 8e2 + 8e3): adjacent pair
    (i, i+1) is estimated on rank i mod W, the 96-byte records are all-gathered (the only collective), every
    rank composes the same H(0 <- i) and canvas geometry and renders its own band of canvas rows straight into
    a host canvas shared by the ranks of the node.  Every rank holds all input images (they come from the
    host).  Returns the panorama (a view of the shared canvas, the same memory on every rank; copy it if it
    has to outlive the next call) or None, plus the gathered per-pair records.  With a CUDA `device` and equally
    sized images the inputs are replicated over NVLink (replicate_images; `replicate=False` or PANO_CHAIN_NVLINK=0
    makes every rank upload all images itself, the variant profiles/r02_chain_*gpu.json was measured with)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if replicate is None:
        replicate = os.environ.get("PANO_CHAIN_NVLINK", "1") != "0"
    n_pairs = len(images) - 1
    mine = shard_pairs(n_pairs, rank, world)
    host_images = images
    if device is not None and replicate and len({np.asarray(im).shape for im in images}) == 1:
        images = replicate_images(images, device, group)
    res = [engine.pairHomography(images[i], images[i + 1]) for i in mine]
    allr = all_gather_results(pack_results(mine, res), n_pairs, device=device, group=group)
    pair_H = [allr[i, :9].reshape(3, 3).copy() if int(allr[i, 9]) == 0 else None for i in range(n_pairs)]
    Hs = engine.composeChain(pair_H)
    sizes = [(np.asarray(im).shape[1], np.asarray(im).shape[0]) for im in host_images]
    ok, geom, T = engine.chainGeometry(sizes, Hs)
    if not ok:
        return None, allr
    cw, ch = geom[0], geom[1]
    shared = SharedCanvas.get(ch * cw * 3, group)
    canvas = shared.buf[:ch * cw * 3].reshape(ch, cw, 3)
    y0, bh = band_rows(ch, rank, world)
    if bh > 0:
        engine.renderChainBand(images, Hs, geom, T, y0, bh, out=canvas[y0:y0 + bh])
    dist.barrier(group=group)            # every band has landed
    return canvas, allr

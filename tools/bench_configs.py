"""The other BASELINE.json configs as bench workloads (bench.py --workload c1|c2|c3|chain).

c1     images/mountain mountain1.jpg + mountain2.jpg (4156x3117), two-image stitch: engine next to the CPU reference
c2     images/oilseed, 4 images folded in sorted order (seed 1), scored against oilseed-ref.jpg with the evaluator
c3     the synthetic 3840x2160 pair of the golden fixtures (numpy generator, seed 267): single-pair latency
chain  config 4: synthetic 8-image 24 MP strip, adjacent pairs sharded across the GPUs, canvas bands per GPU written
       straight into a shared pinned host canvas (strong scaling: one panorama, N GPUs)
Each prints ONE JSON line.  The sample photographs are the reference's data files and are not committed; they are
read from baseline/_ref/images (git-ignored, shipped to the GPU box).  The CPU side is the reference's own code
(oracle/_ref) timed on this box's host cores; where a full serial run would take many minutes (c2: 18 min) the
figure recorded when the golden vectors were generated (tests/golden/ref_runs.json, build container) is quoted and
labelled as such.
"""
import hashlib
import importlib
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"
IMG = os.path.join(ROOT, "baseline", "_ref", "images")
SEED = 12345


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _cpu_model():
    for l in open("/proc/cpuinfo"):
        if l.startswith("model name"):
            return l.split(":", 1)[1].strip()
    return "unknown"


def _golden():
    f = os.path.join(ROOT, "tests", "golden", "ref_runs.json")
    return json.load(open(f)) if os.path.exists(f) else {}


def _time_pair(eng, torch, left, right, reps):
    """engine timings of one pair: resident (device events) and end to end (host arrays in, host canvas out)"""
    L, R = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    dev = []
    for i in range(reps + 2):
        _, r = eng.stitchTwoImages(L, R, fetch=False)
        if i >= 2:
            dev.append(r["ms"]["total"])
    lh, rh = torch.from_numpy(left).pin_memory().numpy(), torch.from_numpy(right).pin_memory().numpy()
    e2e = []
    canvas = None
    for i in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        canvas, r = eng.stitchTwoImages(lh, rh)
        dt = (time.perf_counter() - t0) * 1000
        if i >= 1:
            e2e.append(dt)
    return r, canvas, statistics.median(dev), statistics.median(e2e)


def _ref_pair(left, right, seed, serial=True):
    from oracle import ref as refmod
    out = {}
    ncpu = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(ncpu)
    if refmod.available("omp"):
        R = refmod.Reference("omp")
        t0 = time.perf_counter(); s = R.stitch_pair(left, right, seed=seed); dt = time.perf_counter() - t0
        out["openmp_reference"] = {"seconds": dt, "threads": R.num_threads(), "stage_ms": s["times_ms"], "status": s["status"],
                                   "what": "src/openmp/main.cpp unmodified, -O2 -fopenmp, cvshim"}
    if serial and refmod.available(""):
        S = refmod.Reference("")
        t0 = time.perf_counter(); s = S.stitch_pair(left, right, seed=seed); dt = time.perf_counter() - t0
        out["serial_reference"] = {"seconds": dt, "threads": 1, "stage_ms": s["times_ms"], "status": s["status"],
                                   "canvas_sha": sha(s["canvas"]) if s["canvas"] is not None else None,
                                   "what": "src/serial/main.cpp unmodified, -O2, cvshim"}
    return out


def run_pair_config(a, name):
    import torch
    import cv2
    pkg = importlib.import_module(PKG)
    gold = _golden()
    if name == "c1":
        left = cv2.imread(os.path.join(IMG, "mountain", "mountain1.jpg"))
        right = cv2.imread(os.path.join(IMG, "mountain", "mountain2.jpg"))
        g = gold.get("c1_mountain", {})
        what = "BASELINE config 1: images/mountain mountain1.jpg + mountain2.jpg (4156x3117 each), decoded by cv2.imread"
    else:
        synth = importlib.import_module(PKG + ".synth")
        left, right, _ = synth.make_pair(a.w, a.h, seed=267)
        g = gold.get("c3_pair_4k_seed267", {})
        what = "BASELINE config 3: synthetic %dx%d textured pair, known homography (synth.make_pair seed 267)" % (a.w, a.h)
    eng = pkg.Engine(0, SEED)
    r, canvas, dev_ms, e2e_ms = _time_pair(eng, torch, left, right, 10)
    mp = (left.shape[0] * left.shape[1] + right.shape[0] * right.shape[1]) / 1e6
    line = {"metric": "ms_per_pair", "workload": what, "n_gpus": 1, "input_MP": mp,
            "engine": {"ms_resident_device_median": dev_ms, "ms_end_to_end_wall_median": e2e_ms, "MP_per_s_resident": mp / dev_ms * 1e3,
                       "MP_per_s_end_to_end": mp / e2e_ms * 1e3, "keypoints": [r["kl"], r["kr"]], "matches": r["m"],
                       "inliers": r["best"], "canvas": list(r["canvas"][:2]), "canvas_sha": sha(canvas)},
            "reference_golden": {"canvas_sha": g.get("canvas_sha"), "identical": g.get("canvas_sha") == sha(canvas),
                                 "source": "tests/golden/ref_runs.json (the reference's own code, oracle/_ref)"},
            "cpu_model": _cpu_model(), "host_threads": len(os.sched_getaffinity(0))}
    if not a.no_cpu:
        if os.environ.get("PANO_BENCH_CHILD"):     # (bench.py's other_configs leg: the engine side first, see there)
            print(json.dumps(line, default=float), flush=True)
        line["cpu"] = _ref_pair(left, right, SEED, serial=True)
        for k, v in line["cpu"].items():
            v["MP_per_s"] = mp / v["seconds"]
            v["engine_e2e_speedup"] = v["seconds"] * 1e3 / e2e_ms
    print(json.dumps(line, default=float), flush=True)
    eng.close()


def run_c2(a):
    import torch
    import cv2
    import importlib.util
    pkg = importlib.import_module(PKG)
    gold = _golden().get("c2_oilseed_fold", {})
    ims = [cv2.imread(os.path.join(IMG, "oilseed", "oilseed%d.jpg" % i)) for i in (1, 2, 3, 4)]
    refimg = cv2.imread(os.path.join(IMG, "oilseed-ref.jpg"))
    eng = pkg.Engine(0, 1)
    times = []
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pano, log = eng.stitchAllImages(ims)
        times.append((time.perf_counter() - t0) * 1000)
    spec = importlib.util.spec_from_file_location("evalpano", os.path.join(ROOT, "tools", "evaluate_panorama.py"))
    ev = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ev)
    metrics, _ = ev.compute_metrics(pano, refimg)
    mp = sum(i.shape[0] * i.shape[1] for i in ims) / 1e6
    line = {"metric": "ms_per_panorama", "workload": "BASELINE config 2: images/oilseed, 4 images (2003x1502) folded in sorted order, "
            "seed 1, host images in, host panorama out", "n_gpus": 1, "input_MP": mp,
            "engine": {"ms_end_to_end_wall_median": statistics.median(times[1:]), "canvas": [pano.shape[1], pano.shape[0]],
                       "canvas_sha": sha(pano), "steps": [{k: l[k] for k in ("kl", "kr", "m", "best", "status")} for l in log],
                       "score_vs_oilseed_ref": metrics},
            "reference_golden": {"canvas_sha": gold.get("canvas_sha"), "identical": gold.get("canvas_sha") == sha(pano),
                                 "serial_reference_seconds_build_container": gold.get("ref_seconds"),
                                 "serial_reference_stage_ms_build_container": gold.get("stage_ms"),
                                 "note": "identical output => identical evaluate_panorama score to the reference's"},
            "cpu_model": _cpu_model(), "host_threads": len(os.sched_getaffinity(0))}
    if not a.no_cpu:
        from oracle import ref as refmod
        if refmod.available("omp"):
            os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
            R = refmod.Reference("omp")
            t0 = time.perf_counter(); s = R.stitch_all(ims, seed=1); dt = time.perf_counter() - t0
            line["cpu"] = {"openmp_reference": {"seconds": dt, "threads": R.num_threads(), "status": s["status"], "stage_ms": s["times_ms"],
                                                "engine_e2e_speedup": dt * 1e3 / statistics.median(times[1:]),
                                                "what": "src/openmp/main.cpp unmodified (its RANSAC differs from serial: timing only)"}}
    print(json.dumps(line, default=float), flush=True)
    eng.close()


def run_chain(a):
    """config 4, strong scaling: launched by torchrun with N ranks (or plain python for N = 1)"""
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    pkg = importlib.import_module(PKG)
    pdist = importlib.import_module(PKG + ".dist")
    synth = importlib.import_module(PKG + ".synth")
    n, w, h = 8, 2000, 1500
    views = synth.make_strip(n=n, w=w, h=h, seed=267)
    eng = pkg.Engine(device=local, seed=SEED)
    times = []
    pano = None
    for i in range(a.steps + a.warmup):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        pano, allr = pdist.stitch_chain_distributed(eng, views, device="cuda")
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if i >= a.warmup:
            times.append(float(dt[0]))
    if rank == 0:
        digest = sha(pano)
        single = None
        if world > 1:     # the same panorama on this GPU alone
            ref, _ = eng.stitchChain(views)
            single = bool(ref is not None and ref.shape == pano.shape and np.array_equal(ref, pano))
        mp = n * w * h / 1e6
        t = statistics.median(times)
        print(json.dumps({"metric": "stitched_MP_per_s_chain_panorama", "value": mp / t, "unit": "MP/s", "n_gpus": world,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000 * t, "higher_is_better": True, "scaling": "strong",
                          "workload": "BASELINE config 4: synthetic 8-image 24 MP strip (2000x1500 each, 40 % overlap), chain mode: "
                                      "adjacent pairs sharded over the GPUs, homography records all-gathered (NCCL, 7 x 96 B), "
                                      "each GPU renders its band of canvas rows into a shared pinned host canvas",
                          "canvas": [int(pano.shape[1]), int(pano.shape[0])], "canvas_sha": digest,
                          "identical_to_single_gpu": single, "pairs_ok": int(sum(1 for r in allr if int(r[9]) == 0)),
                          "times_ms": [1000 * x for x in times]}, default=float), flush=True)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


def run(a):
    if a.workload in ("c1", "c3"):
        return run_pair_config(a, a.workload)
    if a.workload == "c2":
        return run_c2(a)
    return run_chain(a)

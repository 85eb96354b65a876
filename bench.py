#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 stitching engine (BASELINE.json metric:
"stitched MP/s and ms per 4K pair (detect+match+RANSAC+warp) at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W                    # engine arm (1 process per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm (rank 0 only)
    python bench.py --workload c1|c2|c3|chain ...                    # the other BASELINE configs (see profiles/)

Default workload = BASELINE config 5: a batch of 256 DISTINCT synthetic 3840x2160 pairs per GPU
(generator synth.make_pair_torch, seeds 1000 + 256 * rank + i; 12.7 GB of input per GPU, far beyond the
126 MB L2).  A step = one pass of the full hot path (detect both images, match, seeded RANSAC, warp +
overlay) over that batch through ONE pano_stitch_batch call.
`value`        whole-job stitched MP/s (input megapixels of all ranks / max-over-ranks device time), inputs
               already resident in HBM, canvases left in HBM.
`e2e`          the same metric through the C-ABI call with HOST (pinned) buffers: H2D of both images and D2H
               of every canvas inside the timed region, distinct host buffers for every pair; also reported as
               a fraction of the pure-copy ceiling measured in the same run (same bytes, no kernels).
`latency`      one pair alone on an idle GPU (BASELINE config 3, "ms per 4K pair"), device and wall time.
`roofline`     the HBM-bound kernel of the step (warp + overlay) timed live with CUDA events around its launch
               inside the C ABI (pano_set_profile), plus every other main kernel timed the same way.
`cpu_baseline` the reference's own code (oracle/_ref: /root/reference/src/openmp/main.cpp and src/serial/main.cpp
               compiled unmodified against oracle/cvshim) on this box's host cores, rank 0 at N = 1 only.
`other_configs` BASELINE configs 1 (images/mountain pair: engine next to the reference's serial and OpenMP code), 2
               (images/oilseed fold + evaluator score) and 4 (8-image strip, chain mode, on this one GPU), each measured
               in a child process at N = 1 (--no-extras skips).  At N > 1 (the scaling runs) config 4 is measured on
               the same N GPUs by a child job of rank 0 after the headline's measurements (chain_config_at_n).
Every number in the line is measured in this run, except those explicitly attributed to a committed capture
under profiles/ (ncu-only metrics such as the tensor-pipe percentage).
"""
import argparse
import os as _os
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # hardware work queues for the batch slots (before CUDA starts)
# NCCL evidence (ranks, transport) goes to stderr so that stdout stays one JSON line
_os.environ.setdefault("NCCL_DEBUG", "INFO")
_os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import importlib
import json
import os
import platform
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "ucb-cs267-parallel-panoramic-image-stitching_b200"
METRIC = "stitched_MP_per_s_4K_pair_detect_match_ransac_warp"
UNIT = "MP/s"
SEED = 12345


def peaks():
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        p = json.load(open(f))
        return p.get("hbm_gbs", 6650.0), "measured copy bandwidth (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_model():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def host_pair(w, h, seed):
    """one pair of the workload on the host (CPU legs): the same generator, evaluated by torch on the CPU"""
    synth = importlib.import_module(PKG + ".synth")
    l, r, _ = synth.make_pair_torch(w, h, seed=seed, device="cpu")
    return l.numpy(), r.numpy()


def workload_name(w, h, P):
    return ("BASELINE config 5: batch of %d distinct synthetic %dx%d textured pairs with known homographies per GPU "
            "(synth.make_pair_torch, seeds 1000 + %d * rank + i)" % (P, w, h, P))


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref), all host threads
# ----------------------------------------------------------------------------------------------
def run_reference(a):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    ncpu = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(ncpu)     # (torchrun exports OMP_NUM_THREADS=1; libgomp reads it at load)
    w, h = a.w, a.h
    from oracle import ref as refmod
    kind = "reference"
    if refmod.available("omp"):
        R = refmod.Reference("omp")
        cores = R.num_threads()
        what = ("the reference's OpenMP pipeline, src/openmp/main.cpp compiled unmodified against oracle/cvshim "
                "(-O2 -fopenmp), %d threads" % cores)

        def run(l, r):
            s = R.stitch_pair(l, r, seed=SEED)   # (its RANSAC samples with std::sample per thread: a timing baseline)
            return s["times_ms"]
    else:   # oracle/_ref did not travel: the OpenCV-free port with OpenMP
        from oracle.oracle import Oracle
        O = Oracle("omp")
        cores, kind = O.num_threads(), "port"
        what = "oracle port (pano_oracle.cpp -O2 -fopenmp), %d threads" % cores

        def run(l, r):
            s = O.stitch_pair(l, r, seed=SEED)
            assert s["status"] == 1
            return s["times_ms"]
    pairs = [host_pair(w, h, 1000 + i) for i in range(2)]
    times, stages = [], None
    for s in range(a.warmup + a.steps):
        l, r = pairs[s % len(pairs)]
        t0 = time.perf_counter()
        stages = run(l, r)
        dt = time.perf_counter() - t0
        if s >= a.warmup:
            times.append(dt)
    mp = 2 * w * h / 1e6
    tot = sum(times)
    val = mp * len(times) / tot
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000 * tot / len(times), "ms_per_pair": 1000 * tot / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+u8", "data": "synthetic",
            # (the engine arm's workload, named identically; what a reference step covers of it is said separately)
            "config": {"workload": workload_name(w, h, a.pairs), "pairs_per_step_per_gpu": a.pairs, "distinct_pairs_per_gpu": a.pairs,
                       "seed": SEED, "reference_step": "a bounded sample of that workload: 1 pair per step (seeds 1000, 1001 alternating)",
                       "pairs_per_reference_step": 1, "cpu_model": cpu_model(), "host_threads": ncpu},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": "%d step(s) of 1 pair; %s" % (len(times), what),
                             "stage_ms_last_step": stages},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line, default=float), flush=True)


# ----------------------------------------------------------------------------------------------
# engine arm
# ----------------------------------------------------------------------------------------------
def pinned_bytes(torch, n):
    return torch.empty(n, dtype=torch.uint8).pin_memory()


def mem_available_bytes():
    try:
        for l in open("/proc/meminfo"):
            if l.startswith("MemAvailable"):
                return int(l.split()[1]) * 1024
    except OSError:
        pass
    return 64 << 30


def copy_ceiling(torch, dist, world, Lh, Rh, Ch, canvas_bytes, Ld, Rd, reps=2):
    """pure copies of the e2e step's bytes (both directions at once, separate streams), all ranks at the same time"""
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    dsrc = torch.empty(max(canvas_bytes), dtype=torch.uint8, device="cuda")
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(len(Lh)):
            with torch.cuda.stream(up):
                Ld[i].copy_(Lh[i], non_blocking=True)
                Rd[i].copy_(Rh[i], non_blocking=True)
            with torch.cuda.stream(down):
                Ch[i][:canvas_bytes[i]].copy_(dsrc[:canvas_bytes[i]], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def run_engine(a):
    import torch
    rank, world, local = dist_env()
    t_start = time.perf_counter()

    def log(msg):      # progress on stderr (stdout carries the one JSON line)
        if rank == 0:
            print("[bench %6.1f s] %s" % (time.perf_counter() - t_start, msg), file=sys.stderr, flush=True)
    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    cpus_before_binding = sorted(os.sched_getaffinity(0))
    numa = importlib.import_module(PKG + ".numa").bind_to_gpu(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ncpu = len(os.sched_getaffinity(0))
    # lanes = host threads driving the batch (each a pipeline over PANO_BATCH_DEPTH + 2 slots); bounded by the cores
    # this rank may use and by the 32 hardware work queues (2 streams per slot + 2 copy streams)
    # (several ranks per box share its cores: twice the rank's share of cores, waits then block instead of spinning)
    lanes = int(os.environ.get("PANO_BATCH_LANES", a.lanes or (max(2, min(10, ncpu - 2)) if world == 1 else max(4, min(10, 2 * ncpu)))))
    os.environ["PANO_BATCH_LANES"] = str(lanes)
    eng = pkg.Engine(device=local, seed=SEED)
    w, h, P = a.w, a.h, a.pairs
    npx = w * h

    # ---- the workload: P distinct pairs, generated on the device ---------------------------------
    t_gen = time.perf_counter()
    Ld, Rd = [], []
    for i in range(P):
        l, r, _ = synth.make_pair_torch(w, h, seed=1000 + rank * P + i, device="cuda")
        Ld.append(l); Rd.append(r)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    log("generated %d pairs per GPU on the device in %.1f s" % (P, t_gen))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pdist = importlib.import_module(PKG + ".dist")
    n_total = world * P
    my_idx = [rank * P + i for i in range(P)]
    est = torch.cuda.ExternalStream(eng.stream_ptr())

    def exchange(raw):
        """the path's only collective: all-gather of the per-pair homography records (96 B each)"""
        if world > 1:
            a = pkg.results_array(raw)
            rec = np.zeros((len(a), pdist.RECORD), np.float64)
            rec[:, :9] = a["H"]; rec[:, 9] = a["status"]; rec[:, 10] = a["best_inliers"]
            rec[:, 11] = rank * len(a) + np.arange(len(a))
            return pdist.all_gather_results(rec, world * len(a), device="cuda")
        return None

    def dicts(raw):
        return [raw[i].as_dict() for i in range(len(raw))]

    def timed(step_fn, steps):
        """K steps bracketed by barrier + synchronize; device time between two events on the engine's stream
        (includes every host gap inside the region)"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(est)
        last = None
        for _ in range(steps):
            last, _ms = step_fn()
            exchange(last)
        e1.record(est)
        e1.synchronize()
        barrier()
        return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0, last

    # ---- resident-input throughput --------------------------------------------------------------
    batch_res = eng.makeBatch(Ld, Rd)      # pointer tables built once: the timed region holds no per-image Python work

    def step_resident():
        return eng.stitchBatch(batch=batch_res, raw=True)

    for _ in range(a.warmup):
        res, _ = step_resident()
        exchange(res)
    bad = [r["status_name"] for r in dicts(res) if r["status"] != 0]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = eng.kernel_launches()
    ms_dev, wall_ms, res = timed(step_resident, a.steps)
    res = dicts(res)
    launches = eng.kernel_launches() - n0
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end (host buffers, distinct for every pair) ---------------------------------------
    canvas_bytes = [3 * r["canvas"][0] * r["canvas"][1] for r in res]
    cap = max(canvas_bytes) + (1 << 20)
    per_pair = 2 * 3 * npx + cap
    budget = int(0.45 * mem_available_bytes() / max(world, 1))
    # pairs of the e2e leg: all P on one GPU (unless host memory is short); with several ranks on one box the leg is
    # bound by the box's shared host memory system, and 128 pairs per GPU already run for seconds per step
    Pe = max(8, min(P if world == 1 else min(P, 128), budget // per_pair))
    log("resident leg done (%.1f ms/step); pinning %.1f GB of host memory for %d e2e pairs" % (ms_dev / a.steps, Pe * per_pair / 1e9, Pe))
    t_pin = time.perf_counter()
    pool = pinned_bytes(torch, Pe * per_pair)
    t_pin = time.perf_counter() - t_pin
    Lh, Rh, Ch = [], [], []
    for i in range(Pe):
        o = i * per_pair
        lh = pool[o:o + 3 * npx].view(h, w, 3); rh = pool[o + 3 * npx:o + 6 * npx].view(h, w, 3)
        lh.copy_(Ld[i]); rh.copy_(Rd[i])
        Lh.append(lh); Rh.append(rh); Ch.append(pool[o + 6 * npx:o + per_pair])
    torch.cuda.synchronize()
    Lnp, Rnp, Cnp = [t.numpy() for t in Lh], [t.numpy() for t in Rh], [t.numpy() for t in Ch]

    batch_e2e = eng.makeBatch(Lnp, Rnp, canvases_out=Cnp)

    def step_e2e():
        return eng.stitchBatch(batch=batch_e2e, raw=True)

    for _ in range(min(a.warmup, 2)):
        step_e2e()
    ms_e2e, _, res_e = timed(step_e2e, a.steps)
    log("e2e leg done (%.1f ms/step)" % (ms_e2e / a.steps))
    res_e = dicts(res_e)
    d2h = sum(3 * r["canvas"][0] * r["canvas"][1] for r in res_e)
    h2d = Pe * 2 * 3 * npx
    # the e2e canvases are the same bytes the resident run produced (spot check: pair 0 against a fresh device canvas)
    t_ceiling = copy_ceiling(torch, dist, world, Lh, Rh, Ch, canvas_bytes[:Pe], Ld, Rd)
    canvas_px = sum(canvas_bytes) // 3       # canvas pixels written per resident step on this rank

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e, wall_ms, t_ceiling], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, wall_ms, t_ceiling = [float(x) for x in t]
        lt = torch.tensor([launches, len(bad), h2d, d2h, canvas_px], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches, nbad, h2d, d2h, canvas_px = [int(x) for x in lt]      # (whole job, like `value`)
    else:
        nbad = len(bad)
    mp_pair = 2 * npx / 1e6
    value = world * P * mp_pair * a.steps / (ms_dev / 1000.0)
    e2e_val = world * Pe * mp_pair * a.steps / (ms_e2e / 1000.0)
    ceiling_val = world * Pe * mp_pair / t_ceiling

    # ---- single-pair latency (rank 0; the other ranks idle at the barrier below) ---------------------
    latency = None
    if rank == 0:
        dev_ms, wall = [], []
        for i in range(3 + 20):
            L, R = Ld[i % P], Rd[i % P]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, r = eng.stitchTwoImages(L, R, fetch=False)
            dt = (time.perf_counter() - t0) * 1000.0
            if i >= 3:
                dev_ms.append(r["ms"]["total"]); wall.append(dt)
        latency = {"ms_per_pair_device_median": statistics.median(dev_ms), "ms_per_pair_wall_median": statistics.median(wall),
                   "ms_per_pair_device_min": min(dev_ms), "pairs": 20,
                   "how": "one pano_stitch_pair call at a time on an otherwise idle GPU, resident inputs, 20 distinct pairs "
                          "after 3 warm-ups; device = CUDA events inside the C ABI, wall = host clock around the call"}

    line = None
    if rank == 0:
        hbm, how = peaks()
        r0 = res[0]
        cw, ch = r0["canvas"][0], r0["canvas"][1]
        # ---- per-kernel device times, measured live: events around the launches (pano_set_profile) ----
        eng.set_profile(True)
        reps = 9
        for i in range(reps):
            eng.stitchTwoImages(Ld[0], Rd[0], fetch=False)
        prof = eng.get_profile()
        eng.set_profile(False)
        k_ms = {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items()}
        k_per_pair = {k: v[0] / reps for k, v in prof.items()}
        mb = {}
        f = os.path.join(ROOT, "profiles", "r02_microbench.json")
        if os.path.exists(f):
            mb = json.load(open(f))
        ncu = {}
        f = os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")
        if os.path.exists(f):
            ncu = json.load(open(f))
        alg_warp = 3 * (2 * npx + cw * ch)
        ach = alg_warp / (k_ms["warp"] / 1e3) / 1e9
        fp64_peak = (mb.get("fp64_instr_per_s") or {}).get("dmul_dadd_pair")
        harris_fp64 = 157.0 * npx * (32 * 64) / (30 * 62)       # incl. the fused kernel's 1-px response halo
        match_ops = 2.0 * r0["kr"] * r0["kl"] * 75
        roofline = {
            "bound": "hbm", "kernel": "warp_quad_kernel (warp.cu)", "achieved": ach, "peak": hbm, "unit": "GB/s",
            "frac": ach / hbm, "traffic": (ncu.get("warp_quad_kernel") or {}).get("dram_bytes"),
            "peak_source": how, "kernel_ms": k_ms["warp"], "algorithmic_bytes": alg_warp,
            "how": "CUDA events on the engine's stream right around the launch, inside pano_stitch_pair, mean of %d "
                   "single-pair runs (the RANSAC kernels before it have flushed the sources from L2)" % reps,
            # the kernel's second ceiling: it is bound by instruction issue, not by bytes
            "issue": issue_fraction((ncu.get("warp_quad_kernel") or {}).get("warp_instructions"), k_ms["warp"], clocks),
            "other_kernels": {
                "harris_fused_kernel": {"ms_per_image": k_ms["harris_fused"], "bound": "fp64 pipe",
                                        "fp64_instr": harris_fp64,
                                        "achieved_fp64_instr_per_s": harris_fp64 / (k_ms["harris_fused"] / 1e3),
                                        "peak_fp64_instr_per_s": fp64_peak,
                                        "frac_of_fp64_peak": (harris_fp64 / (k_ms["harris_fused"] / 1e3) / fp64_peak) if fp64_peak else None,
                                        "peak_source": "tools/microbench.cu on this GPU model (profiles/r02_microbench.json)",
                                        "hbm_frac": 3 * npx / (k_ms["harris_fused"] / 1e3) / 1e9 / hbm},
                "match_tc_kernel": {"ms": k_ms["match_tc"], "bound": "tensor (int8 tcgen05) / epilogue",
                                    "pairs": r0["kr"] * r0["kl"], "achieved_TOPs": match_ops / (k_ms["match_tc"] / 1e3) / 1e12,
                                    **tensor_fraction(r0["kr"], r0["kl"], k_ms["match_tc"], mb, clocks),
                                    "tensor_pipe_active_pct_ncu": (ncu.get("match_tc_kernel") or {}).get("tensor_pipe_active_pct"),
                                    "ncu_source": "profiles/r02_ncu_metrics.json (ncu --set full capture of this kernel)"},
                "replay (all kernels of the shuffle replay)": replay_floor(eng, r0["m"], k_per_pair["replay"], mb, P,
                                                                           ms_dev / (a.steps * P)),
                "dlt_kernel": {"ms": k_ms["dlt"], "bound": "latency (136 dependent Jacobi rotations per hypothesis)"},
                "score_kernel": {"ms": k_ms["score"], "bound": "fp64 pipe"},
                "descriptor_gather": {"ms_per_image": k_ms["descriptor_gather"]},
                "scan_scatter": {"ms_per_image": k_ms["scan_scatter"]}},
            "kernel_ms_sum_per_pair": sum(k_per_pair.values())}
        # ---- CPU baseline: the reference's own code on this box (bounded: one pair, OpenMP build, all threads;
        # the serial build is estimated from a sample of its stages) ---------------------------------
        cpu = None
        if not a.no_cpu and world == 1:
            log("kernel profile done; timing the reference's CPU code on one pair of the workload")
            try:
                cpu = cpu_baseline(a, Ld[0].cpu().numpy(), Rd[0].cpu().numpy(), res[0])
            except Exception as e:      # the line is still worth printing
                cpu = {"value": None, "unit": UNIT, "error": "%s: %s" % (type(e).__name__, e)}
        other = None
        if world == 1 and not a.no_extras:
            log("timing BASELINE configs 1 and 2 in child processes (--no-extras skips them)")
            try:
                other = other_configs(a)
            except Exception as e:
                other = {"error": "%s: %s" % (type(e).__name__, e)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_dev / a.steps, "ms_per_pair": ms_dev / (a.steps * P), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64+u8", "data": "synthetic",
                "config": {"workload": workload_name(w, h, P), "pairs_per_step_per_gpu": P, "distinct_pairs_per_gpu": P,
                           "seed": SEED, "l2_policy": "inputs %.1f GB per GPU, every pair distinct: nothing is re-read from L2" % (P * 6 * npx / 1e9),
                           "lanes": lanes, "pipeline_depth": int(os.environ.get("PANO_BATCH_DEPTH", "1")),
                           "keypoints_pair0": [r0["kl"], r0["kr"]], "matches_pair0": r0["m"], "inliers_pair0": r0["best"],
                           "matches_min_max": [min(r["m"] for r in res), max(r["m"] for r in res)],
                           "failed_pairs": nbad, "parallelism": "pairs sharded over %d GPU(s), no data-path collective" % world,
                           "cpu_model": cpu_model(), "host_threads_this_rank": ncpu, "numa": numa,
                           "generation_s": round(t_gen, 1), "pinned_alloc_s": round(t_pin, 1)},
                "clocks": clocks, "gpu_launches": launches, "gpu_launches_per_pair": launches / (a.steps * P * world),
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_pair": ms_e2e / (a.steps * Pe), "pairs_per_step_per_gpu": Pe,
                        "copy_ceiling": {"value": ceiling_val, "unit": UNIT, "frac": e2e_val / ceiling_val,
                                         "how": "the same H2D / D2H copies alone (both directions at once, all ranks "
                                                "concurrently), best of 2"}},
                "canvas_MP_per_s": canvas_px / 1e6 * a.steps / (ms_dev / 1000.0),
                "latency": latency, "wall_ms_per_step": wall_ms / a.steps,
                "collective": ("all_gather of %d x 96 B homography records per step (NCCL)" % n_total) if world > 1 else "none (single GPU)",
                "roofline": roofline, "cpu_baseline": cpu, "other_configs": other}
    if world > 1 and not a.no_extras:
        # BASELINE config 4 at this GPU count: every measurement of the headline is complete by now
        chain = chain_config_at_n(a, dist, rank, world, cpus_before_binding, log)
        if rank == 0:
            line["other_configs"] = {"chain": chain}
    barrier()
    if rank == 0:
        print(json.dumps(line, default=float), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


TORCHRUN_ENV = ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "GROUP_WORLD_SIZE", "ROLE_RANK",
                "ROLE_WORLD_SIZE", "ROLE_NAME", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS", "PANO_BATCH_LANES")


def chain_child_command(world, cpus, port):
    """`bench.py --workload chain` under its own torchrun with `world` ranks; the bootstrap restores the CPU set this
    rank had before it bound itself next to its GPU (a child inherits the affinity of its parent)"""
    boot = ("import os, sys; os.sched_setaffinity(0, {%s}); os.execv(sys.executable, [sys.executable] + sys.argv[1:])"
            % ", ".join(str(c) for c in cpus))
    return [sys.executable, "-c", boot, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
            "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__), "--workload", "chain",
            "--steps", "5", "--warmup", "2"]


def free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def descendants(pid):
    """every live descendant of `pid` (children of children included), from /proc"""
    kids = {}
    for d in os.listdir("/proc"):
        if d.isdigit():
            try:
                with open("/proc/%s/stat" % d) as f:
                    st = f.read()
                kids.setdefault(int(st[st.rindex(")") + 2:].split()[1]), []).append(int(d))
            except (OSError, ValueError, IndexError):
                pass
    out, todo = [], [pid]
    while todo:
        for k in kids.get(todo.pop(), []):
            out.append(k)
            todo.append(k)
    return out


def run_child_group(cmd, env, timeout):
    """runs `cmd` and returns (exit code or None on timeout, stdout, stderr).  On timeout the child AND every descendant
    is killed (torchrun starts its workers in sessions of their own: a process-group kill would miss them), so that no
    rank of a stuck child job is left on a GPU."""
    import signal
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, start_new_session=True)
    try:
        so, se = p.communicate(timeout=timeout)
        return p.returncode, so, se
    except subprocess.TimeoutExpired:
        victims = descendants(p.pid) + [p.pid]
        for sig, grace in ((signal.SIGTERM, 3.0), (signal.SIGKILL, 0.0)):
            victims = sorted(set(victims + descendants(p.pid)))
            for v in victims:
                try:
                    os.kill(v, sig)
                except OSError:
                    pass
            t_end = time.perf_counter() + grace
            while time.perf_counter() < t_end and p.poll() is None:
                time.sleep(0.1)
        try:
            so, se = p.communicate(timeout=20)
        except subprocess.TimeoutExpired:      # (a survivor still holds the pipes: give the output up, not the run)
            so, se = "", "output of the killed child job abandoned"
        return None, so, se


def chain_config_at_n(a, dist, rank, world, cpus, log=lambda m: None, make_cmd=chain_child_command, store=None):
    """N > 1 (the driver's scaling runs): BASELINE config 4 - the 8-image 24 MP strip in chain mode, pairs and canvas
    bands sharded over the same N GPUs - measured by rank 0 as a CHILD job (`bench.py --workload chain` under its own
    torchrun, the command profiles/r02_chain_*gpu.json was measured with; per-rank uploads) once the headline's
    measurements are complete.  The other ranks wait on the rendezvous store (a host-side wait: no barrier kernel
    spins on their GPUs meanwhile).  A child that fails or runs into the limit leaves an error note."""
    import datetime
    key = "pano_bench_chain_leg_done"
    try:
        store = store or dist.distributed_c10d._get_default_store()
    except Exception as e:
        return {"error": "no rendezvous store: %s" % e} if rank == 0 else None
    if rank != 0:
        try:
            store.wait([key], datetime.timedelta(seconds=a.extras_timeout + 480))
        except Exception:
            pass
        return None
    out = None
    try:
        log("timing BASELINE config 4 (chain mode) on %d GPUs in a child job" % world)
        env = {k: v for k, v in os.environ.items() if k not in TORCHRUN_ENV and not k.startswith("TORCHELASTIC")}
        env.update(PANO_CHAIN_NVLINK="0", PANO_BENCH_CHILD="1", NCCL_DEBUG="WARN")
        t0 = time.perf_counter()
        limit = max(a.extras_timeout, 180.0) if a.extras_timeout >= 60.0 else a.extras_timeout   # (N torch imports, NCCL start-up)
        rc, so, se = run_child_group(make_cmd(world, cpus, free_port()), env, limit)
        lines = [l for l in (so or "").splitlines() if l.startswith("{") and l.rstrip().endswith("}")]
        if lines and (rc == 0 or rc is None):
            out = json.loads(lines[-1])
            out["child_seconds"] = round(time.perf_counter() - t0, 1)
            if rc is None:      # measured and printed, then stuck on its way out
                out["note"] = "the child job had printed its line but had not exited after %.0f s and was killed" % limit
        else:
            out = {"error": ("child job exit code %s" % rc if rc is not None else "child job killed after %.0f s" % limit)
                   + ": " + (se or "")[-300:]}
    except Exception as e:
        out = {"error": "%s: %s" % (type(e).__name__, e)}
    finally:
        try:
            store.set(key, "1")
        except Exception:
            pass
    return out


def issue_fraction(warp_instructions, kernel_ms, clocks):
    """warp instructions of one launch (from the committed ncu capture of the same kernel on the same input size) against
    the issue ceiling of the GPU - 148 SMs x 4 schedulers x 1 instruction per clock at the SM clock sampled in this run -
    over the kernel's event time measured in this run"""
    try:
        mhz = float((clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz"))
        floor_ms = float(warp_instructions) / (148 * 4 * mhz * 1e6) * 1e3
        return {"warp_instructions_ncu": warp_instructions, "issue_floor_ms": floor_ms, "frac_of_issue_peak": floor_ms / kernel_ms,
                "source": "instruction count: profiles/r02_ncu_metrics.json (ncu capture of this kernel on a 4K pair of the same generator: "
                          "approximate for this pair's canvas); "
                          "time and clock: this run"}
    except Exception as e:
        return {"frac_of_issue_peak": None, "note": "not computed: %s" % e}


def tensor_fraction(nq, nt, kernel_ms, mb, clocks):
    """The matcher's MMA work as issued (128 x 128 tiles, K = 75 padded to 128, int8) per second of its event time,
    against the tcgen05 int8 rate tools/microbench.cu measured per SM and clock (profiles/r02_microbench.json) x 148 SMs
    x the SM clock sampled during this run."""
    try:
        per_clk = max(e["mac_per_clk_per_sm"] for e in mb["tcgen05_mma"] if e["kind"] == "i8")
        mhz = float((clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz"))
        macs = (-(-nq // 128) * 128) * (-(-nt // 128) * 128) * 128.0
        peak = per_clk * 148 * mhz * 1e6
        ach = macs / (kernel_ms / 1e3)
        return {"mma_macs_issued": macs, "achieved_mac_per_s": ach, "peak_mac_per_s": peak, "frac_of_tensor_peak": ach / peak,
                "peak_source": "tools/microbench.cu (%d int8 MAC/clk/SM) x 148 SMs x %.0f MHz sampled in this run" % (per_clk, mhz)}
    except Exception as e:
        return {"frac_of_tensor_peak": None, "tensor_peak_note": "not computed: %s" % e}


def replay_floor(eng, m, replay_ms, mb, batch_pairs, batch_ms_per_pair):
    """The shuffle replay against its operation-count floor: pass 1 evaluates `cells` tests of 3 integer-ALU instructions
    (pano_replay_work_estimate: the plan for this match count), the cell kernel's measured rate on an otherwise idle GPU
    is tools/microbench.cu's (profiles/r02_microbench.json).  Single-pair plan (chunks of 50 000 candidate walks) against
    the replay's event time of the profiled pair; throughput-mode plan (4000 per chunk) against the whole step."""
    out = {"ms_per_pair": replay_ms, "bound": "integer ALU + dependent phases"}
    try:
        rate = float(mb.get("replay_cells_per_s"))
        w1, wb = eng.replayWork(m, 1000, 0.0), eng.replayWork(m, 1000, 4000.0)
        f1, fb = w1["cells"] / rate * 1e3, wb["cells"] / rate * 1e3
        out.update({"cells_per_pair": w1["cells"], "chunks": w1["chunks"], "floor_ms": f1, "floor_over_measured": f1 / replay_ms,
                    "cell_rate_per_s": rate, "cell_rate_source": "tools/microbench.cu (profiles/r02_microbench.json)",
                    "throughput_mode": {"cells_per_pair": wb["cells"], "chunks": wb["chunks"], "floor_ms_per_pair": fb,
                                        "share_of_ms_per_pair": fb / batch_ms_per_pair}})
    except Exception as e:
        out["floor"] = "not computed: %s" % e
    return out


def other_configs(a, runner=subprocess.run):
    """BASELINE configs 1 (mountain pair, beside the reference's CPU code) and 2 (oilseed fold, scored with the evaluator)
    measured in the same run: each in a CHILD process (`bench.py --workload c1|c2`, tools/bench_configs.py), so that
    nothing they do - a missing image, an error, a crash - can touch the headline line; their JSON lines are embedded."""
    out = {}
    # (chain = BASELINE config 4 on this one GPU; its 2 / 4 / 8-GPU lines are profiles/r02_chain_*gpu.json, measured with
    # the same command under torchrun and with the per-rank uploads PANO_CHAIN_NVLINK=0 selects)
    for name, extra in (("c1", []), ("c2", ["--no-cpu"]), ("chain", ["--steps", "5", "--warmup", "2"])):
        try:
            t0 = time.perf_counter()
            # PANO_BENCH_CHILD: the child prints its line once the engine side is measured and again with the CPU legs, so
            # that a CPU leg running into the time limit on a slow host costs that leg only
            p = runner([sys.executable, os.path.abspath(__file__), "--workload", name] + extra, capture_output=True, text=True,
                       timeout=a.extras_timeout, env=dict(os.environ, PANO_BENCH_CHILD="1", PANO_CHAIN_NVLINK="0"))
            lines = [l for l in (p.stdout or "").splitlines() if l.startswith("{")]
            if p.returncode == 0 and lines:
                out[name] = json.loads(lines[-1])
                out[name]["child_seconds"] = round(time.perf_counter() - t0, 1)
            else:
                out[name] = {"error": "child exit code %s: %s" % (p.returncode, (p.stderr or "")[-300:])}
        except subprocess.TimeoutExpired as e:
            so = e.stdout.decode("utf-8", "replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
            lines = [l for l in so.splitlines() if l.startswith("{") and l.rstrip().endswith("}")]
            try:
                out[name] = json.loads(lines[-1])
                out[name]["cpu_legs"] = "not finished within %.0f s (engine side complete)" % a.extras_timeout
            except Exception:
                out[name] = {"error": "TimeoutExpired: %s" % e}
        except Exception as e:      # unparsable output: the headline does not depend on it
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
    return out


def cpu_baseline(a, l, r, eng_res):
    """The reference's own code on this box's host cores.  Primary figure: its OpenMP pipeline on all threads, one
    full pair of the workload.  Also: its serial pipeline (1 core), estimated from a sample - detection of both
    images and RANSAC in full, the matcher on every 16th query keypoint (its cost is linear in the queries)."""
    ncpu = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(ncpu)
    from oracle import ref as refmod
    mp = 2 * a.w * a.h / 1e6
    if not refmod.available("omp"):
        from oracle.oracle import Oracle
        O = Oracle("omp")
        t0 = time.perf_counter()
        o = O.stitch_pair(l, r, seed=SEED)
        dt = time.perf_counter() - t0
        return {"value": mp / dt, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
                "sample": "1 pair, oracle port -O2 -fopenmp (oracle/_ref not present on this box), %.2f s" % dt}
    R = refmod.Reference("omp")
    t0 = time.perf_counter()
    s = R.stitch_pair(l, r, seed=SEED)
    dt = time.perf_counter() - t0
    out = {"value": mp / dt, "unit": UNIT, "cores": R.num_threads(), "kind": "reference", "cpu_model": cpu_model(),
           "sample": "1 full pair of the workload (pair 0), the reference's OpenMP pipeline (src/openmp/main.cpp, "
                     "unmodified, -O2 -fopenmp, cvshim), %.2f s" % dt,
           "stage_ms": s["times_ms"]}
    try:    # what the reference's own CMake flags produce (no build type = -O0; ref: CMakeLists.txt), same pair, same threads
        if refmod.available("omp_O0"):
            R0 = refmod.Reference("omp_O0")
            t0 = time.perf_counter()
            s0 = R0.stitch_pair(l, r, seed=SEED)
            dt0 = time.perf_counter() - t0
            out["reference_O0_build"] = {"value": mp / dt0, "unit": UNIT, "cores": R0.num_threads(), "status": s0["status"],
                                         "sample": "the same pair, src/openmp/main.cpp compiled -O0 -fopenmp (the reference's "
                                                   "default build), %.2f s" % dt0}
    except Exception as e:
        out["reference_O0_build"] = {"error": str(e)}
    try:
        S = refmod.Reference("")
        t0 = time.perf_counter(); kl = S.detect(l); kr = S.detect(r); t_det = time.perf_counter() - t0
        sub = kr[::16]
        t0 = time.perf_counter(); S.match(sub, kl, r, l); t_m = (time.perf_counter() - t0) * len(kr) / max(len(sub), 1)
        m = R.match(kr, kl, r, l)          # (full match list from the OpenMP build, only to feed RANSAC)
        t0 = time.perf_counter(); H = S.ransac(kr, kl, m, seed=SEED); t_r = time.perf_counter() - t0
        same = H is not None and np.array_equal(np.ascontiguousarray(H).view(np.uint64), eng_res["H"].view(np.uint64))
        out["serial_reference"] = {"value_estimate": mp / (t_det + t_m + t_r), "cores": 1,
                                   "stage_s": {"detect_both": t_det, "match_extrapolated": t_m, "ransac": t_r},
                                   "sample": "src/serial/main.cpp unmodified (-O2, cvshim): detection and RANSAC in full, matcher "
                                             "on every 16th query keypoint x16; warp + overlay (~1 s) not included",
                                   "H_bit_identical_to_engine": bool(same)}
    except Exception as e:   # the serial estimate is optional
        out["serial_reference"] = {"error": str(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c1", "c2", "c3", "chain"])
    ap.add_argument("--pairs", type=int, default=256, help="distinct 4K pairs per GPU per step (BASELINE config 5: 256)")
    ap.add_argument("--lanes", type=int, default=0, help="batch lanes (host threads) per GPU; 0 = from the core count")
    ap.add_argument("--size", default="3840x2160")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the child runs of BASELINE configs 1 / 2 / 4 (other_configs)")
    ap.add_argument("--extras-timeout", type=float, default=120.0, help="seconds per child run of other_configs")
    a = ap.parse_args()
    a.w, a.h = [int(v) for v in a.size.split("x")]
    if a.workload != "c5":
        other = importlib.import_module("tools.bench_configs")
        return other.run(a)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)


if __name__ == "__main__":
    main()

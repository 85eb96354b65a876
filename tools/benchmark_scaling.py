#!/usr/bin/env python
"""Scaling / comparison harness: the idea of the reference's benchmark_panorama.py
(strong + weak scaling CSV) and benchmark_serial_parallel.py (serial vs parallel per dataset),
for GPU counts instead of OpenMP thread counts and without matplotlib.

  python tools/benchmark_scaling.py gpus --counts 1 2 4 8 [--steps 10 --warmup 3] [--csv scaling.csv]
      runs bench.py at every N (torchrun for N > 1, 127.0.0.1 rendezvous), plus the CPU reference arm once,
      and writes one CSV row per N: value, e2e, ms per pair, weak-scaling efficiency vs N = 1.
  python tools/benchmark_scaling.py datasets --root images [--impls serial openmp gpu] [--csv datasets.csv]
      runs `./pano.sh run <impl> --dir <dataset>` for every sub-directory of --root and parses the stage lines
      the reference's scripts grep ('Image Stitching', 'Total Execution Time'), one CSV row per (dataset, impl).
"""
import argparse
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(n, steps, warmup, impl="engine", port=29611):
    args = ["bench.py", "--gpus", str(n), "--steps", str(steps), "--warmup", str(warmup)]
    if impl != "engine":
        args += ["--impl", impl]
    if n > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
               "--master-addr", "127.0.0.1", "--master-port", str(port)] + args
    else:
        cmd = [sys.executable] + args
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    for line in reversed(p.stdout.splitlines()):
        line = line.strip()
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("no JSON line from %s\n%s" % (" ".join(cmd), p.stderr[-2000:]))


def cmd_gpus(a):
    rows, base = [], None
    ref = None
    if not a.no_reference:
        ref = run_bench(1, max(1, a.steps // 5), 1, impl="reference")
    for n in a.counts:
        d = run_bench(n, a.steps, a.warmup, port=29611 + n)
        if base is None:
            base = d["value"] / d["n_gpus"]
        rows.append({"n_gpus": d["n_gpus"], "value_MPs": round(d["value"], 1), "e2e_MPs": round(d["e2e"]["value"], 1),
                     "ms_per_pair": round(d.get("ms_per_pair", 0.0), 4),
                     "weak_efficiency": round(d["value"] / (base * d["n_gpus"]), 4),
                     "cpu_reference_MPs": round(ref["value"], 2) if ref else "",
                     "e2e_vs_cpu_reference": round(d["e2e"]["value"] / ref["value"], 1) if ref else ""})
        print(rows[-1])
    with open(a.csv, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)
    print("wrote", a.csv)


STAGE = re.compile(r"^(Harris Corner Detection|Harris Corner Matching|RANSAC Homography Estimation|Image Stitching|"
                   r"Total Stitching Process|Total Execution Time)[^:]*:\s*([0-9.]+)\s*ms")


def cmd_datasets(a):
    rows = []
    for name in sorted(os.listdir(a.root)):
        d = os.path.join(a.root, name)
        if not os.path.isdir(d):
            continue
        for impl in a.impls:
            out = os.path.join(a.out_dir, "%s_%s.png" % (name, impl))
            env = dict(os.environ, PANO_SORT_DIR="1")
            p = subprocess.run(["./pano.sh", "run", impl, "--dir", d, "--out", out], cwd=ROOT, capture_output=True,
                               text=True, env=env)
            t = {}
            for line in p.stdout.splitlines():
                m = STAGE.match(line.strip())
                if m:
                    t[m.group(1)] = t.get(m.group(1), 0.0) + float(m.group(2))
            rows.append({"dataset": name, "impl": impl, "ok": int(os.path.exists(out)),
                         "image_stitching_ms": round(t.get("Image Stitching", float("nan")), 3),
                         "total_execution_ms": round(t.get("Total Execution Time", float("nan")), 3)})
            print(rows[-1])
    if rows:
        with open(a.csv, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
            w.writeheader()
            w.writerows(rows)
        print("wrote", a.csv)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    g = sub.add_parser("gpus")
    g.add_argument("--counts", type=int, nargs="+", default=[1, 2, 4, 8])
    g.add_argument("--steps", type=int, default=10)
    g.add_argument("--warmup", type=int, default=3)
    g.add_argument("--csv", default="scaling.csv")
    g.add_argument("--no-reference", action="store_true")
    g.set_defaults(fn=cmd_gpus)
    d = sub.add_parser("datasets")
    d.add_argument("--root", default="images")
    d.add_argument("--impls", nargs="+", default=["serial", "openmp", "gpu"])
    d.add_argument("--out-dir", default="/tmp")
    d.add_argument("--csv", default="datasets.csv")
    d.set_defaults(fn=cmd_datasets)
    a = ap.parse_args()
    a.fn(a)


if __name__ == "__main__":
    main()

"""CPU tier: host-side helpers of the bench / multi-GPU path (no GPU needed)."""
import importlib
import os

import numpy as np

from conftest import PKG, ROOT, load_pkg, load_synth


def test_numa_helpers_parse_and_degrade_gracefully():
    numa = importlib.import_module(PKG + ".numa")
    assert numa._parse_cpulist("0-3,8,10-11") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set() and numa._parse_cpulist(None) == set()
    before = os.sched_getaffinity(0)
    try:
        info = numa.bind_to_gpu(0, 1)            # no GPU here: nothing to bind to, must not raise
        assert info["gpu"] == 0 and info["allowed_cpus"] == len(before)
        info2 = numa.bind_to_gpu(1, 2)           # two ranks, GPU node unknown: even split of the allowed CPUs
        if len(before) >= 2:
            assert info2.get("bound_cpus") == len(sorted(before)[1::2])
    finally:
        os.sched_setaffinity(0, before)


def test_torch_generator_draws_the_same_scene_as_the_numpy_one():
    """bench.py's device-side generator uses the numpy generator's random parameters: same left image up to the last
    bit of float32 rounding, same kind of right image (its +-2 LSB noise comes from another RNG)"""
    synth = load_synth()
    l, r, H = synth.make_pair_torch(640, 360, seed=1003, device="cpu")
    l2, r2, H2 = synth.make_pair(640, 360, seed=1003)
    assert np.allclose(H, H2)
    d = np.abs(l.numpy().astype(np.int16) - l2.astype(np.int16))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    assert np.abs(r.numpy().astype(np.int16) - r2.astype(np.int16)).mean() < 3.0


def test_pair_result_numpy_view_matches_the_c_struct():
    import ctypes as C
    pkg = importlib.import_module(PKG)
    arr = (pkg.PairResult * 3)()
    arr[1].status = 5; arr[1].n_matches = 77; arr[1].best_inliers = 9; arr[1].H[4] = 2.5
    arr[2].canvas.canvas_w = 123; arr[2].ms_total = 1.5
    v = pkg.results_array(arr)
    assert v.shape == (3,) and v["status"][1] == 5 and v["n_matches"][1] == 77 and v["best_inliers"][1] == 9
    assert v["H"][1][4] == 2.5 and v["canvas"]["canvas_w"][2] == 123 and v["ms_total"][2] == 1.5
    assert pkg.PAIR_DTYPE.itemsize == C.sizeof(pkg.PairResult)


def test_stall_summary_tool_reads_the_committed_source_pages():
    """tools/ncu_stalls.py on the committed per-instruction pages: totals are consistent and the kernels'
    signature instructions are where DESIGN.md says they are"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ncu_stalls", os.path.join(ROOT, "tools", "ncu_stalls.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    d = os.path.join(ROOT, "profiles", "r02_ncu_csv")
    m = mod.summarise(os.path.join(d, "r02_match_tc_kernel.source_sass.csv"))
    assert "match_tc_kernel" in m["kernel"] and m["samples"] > 500
    assert m["opcode_executed"].get("IMAD", 0) > 0 and sum(m["stall_reason_pct"].values()) <= 100.5
    w = mod.summarise(os.path.join(d, "r02_warp_quad_kernel.source_sass.csv"))
    assert "warp_quad_kernel" in w["kernel"] and w["opcode_executed"].get("IDP", 0) > 0        # DP4A bilinear
    assert w["hottest"][0]["stall"] == "stall_long_sb"                                       # tap-load latency
    h = mod.summarise(os.path.join(d, "r02_harris_fused_kernel.source_sass.csv"))
    assert h["opcode_executed"]["DMUL"] > 2e7 and h["opcode_executed"]["DADD"] > 2e7           # separately rounded: no DFMA
    assert h["opcode_executed"].get("DFMA", 0) == 0


def test_bench_other_configs_embeds_child_lines_and_survives_failures():
    """bench.py's other_configs leg: the BASELINE config 1 / 2 child runs are embedded when they succeed and reduced to
    an error note when they fail, time out or print garbage - the headline line never depends on them"""
    import importlib.util
    import subprocess
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a = types.SimpleNamespace(extras_timeout=5.0)
    calls = []

    def runner(cmd, **kw):
        calls.append(cmd)
        name = cmd[cmd.index("--workload") + 1]
        if name == "c1":
            return types.SimpleNamespace(returncode=0, stdout='noise\n{"metric": "ms_per_pair", "engine": {"ms": 1.5}}\n', stderr="")
        raise subprocess.TimeoutExpired(cmd, 5.0)
    out = bench.other_configs(a, runner=runner)
    assert out["c1"]["metric"] == "ms_per_pair" and out["c1"]["engine"]["ms"] == 1.5 and "child_seconds" in out["c1"]
    assert "TimeoutExpired" in out["c2"]["error"]
    assert [c[c.index("--workload") + 1] for c in calls] == ["c1", "c2", "chain"] and "--no-cpu" in calls[1]
    # a child whose CPU leg ran into the limit has already printed its engine-side line: that line is kept
    def slow(cmd, **kw):
        assert kw["env"]["PANO_BENCH_CHILD"] == "1"
        raise subprocess.TimeoutExpired(cmd, 5.0, output=b'log\n{"metric": "ms_per_pair", "engine": {"ms": 2.5}}\n{"metric": "trunc')
    part = bench.other_configs(a, runner=slow)
    assert part["c1"]["engine"]["ms"] == 2.5 and "not finished" in part["c1"]["cpu_legs"]
    bad = bench.other_configs(a, runner=lambda cmd, **kw: types.SimpleNamespace(returncode=3, stdout="", stderr="boom"))
    assert "exit code 3" in bad["c1"]["error"] and "boom" in bad["c2"]["error"]



def test_bench_floor_helpers_compute_from_the_plan_and_the_microbenchmarks():
    """bench.py's live fractions: the replay against its operation-count floor (pano_replay_work_estimate x the measured
    cell rate) and the matcher's issued MMA work against the measured tcgen05 int8 rate; missing inputs leave a note"""
    import ctypes as C
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    pkg = load_pkg()
    e = pkg.Engine.__new__(pkg.Engine)
    e.lib = pkg.load_library()
    mb = json.load(open(os.path.join(ROOT, "profiles", "r02_microbench.json")))
    r = bench.replay_floor(e, 10833, 0.84, mb, 256, 0.753)
    assert r["chunks"] == 8 and 1e9 < r["cells_per_pair"] < 4e9 and 0.1 < r["floor_over_measured"] < 0.5
    assert r["throughput_mode"]["cells_per_pair"] < r["cells_per_pair"] and 0 < r["throughput_mode"]["share_of_ms_per_pair"] < 0.5
    assert "not computed" in bench.replay_floor(e, 10833, 0.84, {}, 256, 0.753)["floor"]
    t = bench.tensor_fraction(10833, 10797, 0.02834, mb, {"sm_mhz": 1965.0})
    assert t["mma_macs_issued"] == 10880 * 10880 * 128 and 0.15 < t["frac_of_tensor_peak"] < 0.3
    assert bench.tensor_fraction(10833, 10797, 0.02834, {}, None)["frac_of_tensor_peak"] is None
    i = bench.issue_fraction(32311417.0, 0.0487, {"sm_mhz": 1965.0})
    assert abs(i["issue_floor_ms"] - 0.02778) < 1e-4 and 0.5 < i["frac_of_issue_peak"] < 0.65
    assert bench.issue_fraction(None, 0.0487, None)["frac_of_issue_peak"] is None

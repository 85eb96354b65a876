"""Extracts the handful of ncu metrics bench.py quotes (DRAM bytes, tensor-pipe activity, FP64 pipe, duration) from
`ncu --set full` reports and writes profiles/<tag>_ncu_metrics.json.   usage: ncu_metrics.py TAG report.ncu-rep ..."""
import csv
import io
import json
import os
import subprocess
import sys

# ncu prints every value in a unit of its own choosing (byte / Kbyte / Mbyte, ns / us / ms): normalised to bytes and ns here
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9,
         "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}
WANT = {"gpu__time_duration.sum": "duration_ns", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_active_pct",
        "sm__inst_executed.avg.per_cycle_elapsed": "ipc_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "smsp__inst_executed.sum": "warp_instructions", "lts__t_bytes.sum": "l2_bytes",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct"}


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    out = {}
    for rep in reps:
        p = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(p.stdout)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = d.get("Kernel Name", "?").split("(")[0].replace("void ", "").split("::")[-1].split("<")[0]
            e = {}
            for k, short in WANT.items():
                if k in d and d[k] != "":
                    try:
                        unit = units[hdr.index(k)]
                        e[short] = float(d[k].replace(",", "")) * SCALE.get(unit, 1.0)
                        e[short + "_unit"] = {"byte": "byte", "ns": "ns"}.get(
                            "byte" if unit.endswith("byte") else ("ns" if unit.endswith("s") and unit in SCALE else unit), unit)
                    except ValueError:
                        pass
            if "dram_read" in e and "dram_write" in e:
                e["dram_bytes"] = e["dram_read"] + e["dram_write"]
            e["report"] = os.path.basename(rep)
            out[name] = e
    json.dump(out, open(os.path.join("profiles", tag + "_ncu_metrics.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:3000])


if __name__ == "__main__":
    main()

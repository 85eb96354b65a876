"""Extracts the handful of ncu metrics bench.py quotes (DRAM bytes, tensor-pipe activity, FP64 pipe, duration) from
`ncu --set full` reports and writes profiles/<tag>_ncu_metrics.json.   usage: ncu_metrics.py TAG report.ncu-rep ..."""
import csv
import io
import json
import os
import subprocess
import sys

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
                        e[short] = float(d[k].replace(",", ""))
                        e[short + "_unit"] = units[hdr.index(k)]
                    except ValueError:
                        pass
            if "dram_read" in e and "dram_write" in e:
                e["dram_bytes"] = e["dram_read"] + e["dram_write"]   # (units in *_unit)
            e["report"] = os.path.basename(rep)
            out[name] = e
    json.dump(out, open(os.path.join("profiles", tag + "_ncu_metrics.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:3000])


if __name__ == "__main__":
    main()

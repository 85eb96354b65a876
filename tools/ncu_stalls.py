"""Stall-sample summary of an ncu source page exported as CSV
(`ncu -i X.ncu-rep --page source --csv --print-source sass > X.source_sass.csv`).

    python tools/ncu_stalls.py profiles/r02_ncu_csv/r02_match_tc_kernel.source_sass.csv [--top 12]

Prints the share of warp-state samples per stall reason, per instruction class (opcode) and the hottest
instructions; runs on the CPU (no ncu needed), so the committed CSVs can be re-read by anyone."""
import argparse
import collections
import csv
import json
import re


def load(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    return rows[0][1] if rows[0] and rows[0][0] == "Kernel Name" else "?", hdr, [r for r in rows[h + 1:] if len(r) == len(hdr)]


def num(v):
    try:
        return int(v)
    except ValueError:
        try:
            return int(float(v))
        except ValueError:
            return 0


def opcode(src):
    t = src.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return re.split(r"[.]", t[0])[0] if t else "?"


def summarise(path, top=12):
    name, hdr, data = load(path)
    ix = {k: i for i, k in enumerate(hdr)}
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    total = sum(num(r[ix["# Samples"]]) for r in data)
    by_reason = collections.Counter()
    by_op = collections.Counter()
    exec_by_op = collections.Counter()
    for r in data:
        for s in stalls:
            by_reason[s] += num(r[ix[s]])
        op = opcode(r[ix["Source"]])
        by_op[op] += num(r[ix["# Samples"]])
        exec_by_op[op] += num(r[ix["Instructions Executed"]])
    hot = sorted(data, key=lambda r: -num(r[ix["# Samples"]]))[:top]
    pct = (lambda n: round(100.0 * n / total, 1)) if total else (lambda n: 0.0)
    return {"kernel": name.split("(")[0], "samples": total,
            "warp_instructions": sum(num(r[ix["Instructions Executed"]]) for r in data),
            "stall_reason_pct": {k: pct(v) for k, v in by_reason.most_common(8) if v},
            "opcode_sample_pct": {k: pct(v) for k, v in by_op.most_common(10) if v},
            "opcode_executed": {k: v for k, v in exec_by_op.most_common(10)},
            "hottest": [{"pct": pct(num(r[ix["# Samples"]])), "executed": num(r[ix["Instructions Executed"]]),
                         "sass": " ".join(r[ix["Source"]].split())[:80],
                         "stall": max(stalls, key=lambda s: num(r[ix[s]]))} for r in hot]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv", nargs="+")
    ap.add_argument("--top", type=int, default=12)
    a = ap.parse_args()
    for f in a.csv:
        print(json.dumps(summarise(f, a.top), indent=1))


if __name__ == "__main__":
    main()

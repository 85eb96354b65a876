#!/usr/bin/env bash
# SASS listing per kernel of libpano_b200.so (north_star: "each kernel's design is backed by ... plus a SASS listing").
# Usage: tools/dump_sass.sh [TAG]   -> profiles/sass_<TAG>/<kernel>.sass  + summary.txt (instruction mix per kernel)
set -eu
TAG=${1:-r02}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SO="$ROOT/ucb-cs267-parallel-panoramic-image-stitching_b200/libpano_b200.so"
OUT="$ROOT/profiles/sass_$TAG"
mkdir -p "$OUT"
cuobjdump -sass "$SO" > "$OUT/all.tmp"
python3 - "$OUT" <<'PY'
import collections, os, re, subprocess, sys
out = sys.argv[1]
txt = open(os.path.join(out, "all.tmp")).read()
parts = re.split(r"\n\s*Function : ", txt)
summary = []
for p in parts[1:]:
    mangled, body = p.split("\n", 1)
    try:
        name = subprocess.run(["c++filt", mangled.strip()], capture_output=True, text=True).stdout.strip()
    except Exception:
        name = mangled.strip()
    short = re.sub(r"^void ", "", name)
    short = re.sub(r"pano::\(anonymous namespace\)::", "", short).split("(")[0]
    fname = re.sub(r"[^A-Za-z0-9_]+", "_", short).strip("_")[:80]
    ops = collections.Counter()
    n = 0
    for line in body.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            ops[m.group(1).split(".")[0]] += 1
            n += 1
    # keep address + instruction, drop the hex encodings (two thirds of the bytes)
    lines = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in body.split("\n")]
    lines = [l for l in lines if l.strip() and not re.match(r"^\s*/\* 0x[0-9a-f]+ \*/$", l)]
    open(os.path.join(out, fname + ".sass"), "w").write("Function : %s\n%s\n%s\n" % (mangled, name, "\n".join(lines)))
    key = [k for k in ("UTCIMMA", "UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "DFMA", "DMUL", "DADD", "IDP", "LDG", "STG", "LDS", "IMAD", "VIMNMX3", "MUFU") if ops.get(k)]
    summary.append("%-60s %6d instr  %s" % (short[:60], n, " ".join("%s=%d" % (k, ops[k]) for k in key)))
open(os.path.join(out, "summary.txt"), "w").write("\n".join(sorted(summary)) + "\n")
print("\n".join(sorted(summary)))
PY
rm -f "$OUT/all.tmp"

"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path, skip_frac=0.5):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    data = data[int(len(data) * skip_frac):]      # steady-state part (later repetitions)
    agg = collections.OrderedDict()
    for r in data:
        name = r[ki].split("(")[0].replace("void ", "").replace("pano::<unnamed>::", "").replace("unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-42s n=%4d %10.1f us %6.1f%%  (%.1f us each)" % (k[:42], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
    print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST step."""
import collections
import csv
import re
import sys


def main(path, detail=False):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",", "")) / 1e3, row.get("Grid Size", "")))
    starts = [i for i, x in enumerate(rows) if "cast_kernel" in x[0]]
    last = rows[starts[-1]:] if starts else rows

    def short(n):
        n = re.sub(r"\(.*", "", n)
        return re.sub(r"void |fervit::|\(anonymous namespace\)::", "", n)[:72]
    agg = collections.OrderedDict()
    for n, t, g in last:
        k = short(n)
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"kernels in last step: {len(last)}, total {tot:.1f} us")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:9.1f} us {v[0]:4d}  {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0]:6.1f}  {k}")
    if detail:
        for i, (n, t, g) in enumerate(last):
            print(i, f"{t:8.1f}us", g, short(n))


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) > 2)

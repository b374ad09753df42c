"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the launch sequence of
the LAST call (the launches after the last Philox draw, or all of them).  Usage: launch_summary.py file.csv [--seq]"""
import collections
import csv
import re
import sys


def load(fn):
    rows = []
    with open(fn) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            u = r["Metric Unit"]
            v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
            rows.append((int(r["ID"]), r["Kernel Name"], v))
    return rows


def short(n):
    n = re.sub(r"void corrla::\(anonymous namespace\)::", "", n)
    n = re.sub(r"corrla::\(anonymous namespace\)::", "", n)
    return re.sub(r"\(.*", "", n)


def main():
    fn = sys.argv[1]
    rows = load(fn)
    idx = [i for i, (_, n, _) in enumerate(rows) if "philox_normal" in n]
    call = rows[idx[-1]:] if idx else rows
    tot = sum(v for _, _, v in call)
    print(f"{fn}: {len(rows)} launches captured; last call: {len(call)} launches, {tot / 1e3:.3f} ms (cold-cache, serialised)")
    agg = collections.OrderedDict()
    for _, n, v in call:
        a = agg.setdefault(short(n), [0, 0.0])
        a[0] += 1
        a[1] += v
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v:11.1f} us {100 * v / tot:5.1f}% {c:4d}x  {n[:110]}")
    if "--seq" in sys.argv:
        print("--- sequence")
        for _, n, v in call:
            print(f"{v:10.1f}  {short(n)[:100]}")


if __name__ == "__main__":
    main()

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share.
Usage: python profiles/summarize_launches.py gpurun_out/launches.csv [top_n]"""
import collections
import csv
import sys


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    out = []
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        out.append((r[ki], v))
    return out


def short(name):
    n = name.split("(")[0]
    return n.replace("void ", "").replace("<unnamed>::", "")[:70]


if __name__ == "__main__":
    data = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for name, us in data:
        a = agg[short(name)]
        a[0] += 1
        a[1] += us
        a[2] = max(a[2], us)
    tot = sum(v[1] for v in agg.values())
    print(f"{len(data)} launches, {tot:.1f} us total")
    print(f"{'kernel':72s} {'n':>5s} {'total us':>10s} {'share':>6s} {'avg us':>8s} {'max us':>8s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k:72s} {v[0]:5d} {v[1]:10.1f} {v[1] / tot:6.3f} {v[1] / v[0]:8.1f} {v[2]:8.1f}")

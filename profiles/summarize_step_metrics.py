"""Per-kernel-family summary of an `ncu --metrics ... --csv` pass over consecutive eager training steps (long format CSV):
one period (= one step) of the launch sequence is located, then time, DRAM bytes and tensor-pipe activity are summed per
kernel family. Writes profiles/ncu_metrics.json (read by bench.py: `roofline.traffic`, `pipe_tensor_active_pct`).

    python profiles/summarize_step_metrics.py gpurun_out/r02_step_metrics_r224_tf32.csv [--json profiles/ncu_metrics.json]"""
import collections
import csv
import json
import re
import sys


def short(name):
    m = re.search(r"(?:tc::|<unnamed>::|\(anonymous namespace\)::)?([A-Za-z_0-9]+)(<[^>(]*>)?\(", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


def family(k):
    if k.startswith(("conv_", "gemm_tc", "igemm", "stem_", "gather_subfilters", "zero_class")):
        return "conv family (xsb_conv2d_fwd + dgrad + wgrad)"
    if k.startswith(("col_sums_atomic", "bn_bwd", "pool_bn")):
        return "xsb_bn_bwd_dxsum"
    if k.startswith("bn_apply"):
        return "xsb_bn_apply"
    if k.startswith("bn_finalize"):
        return "xsb_bn_finalize"
    return "other"


def main():
    path = sys.argv[1]
    rows = collections.OrderedDict()
    with open(path) as f:
        rd = csv.reader(l for l in f if l.startswith('"'))
        hdr = next(rd)
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rd:
            d = rows.setdefault(int(r[ix["ID"]]), {"name": short(r[ix["Kernel Name"]]), "grid": r[ix["Grid Size"]]})
            v = float(r[ix["Metric Value"]].replace(",", ""))
            unit = r[ix["Metric Unit"]]
            if unit in ("ns", "nsecond"):
                v /= 1e3
            elif unit in ("ms", "msecond"):
                v *= 1e3
            elif unit in ("Mbyte",):
                v *= 1e6
            elif unit in ("Kbyte",):
                v *= 1e3
            elif unit in ("Gbyte",):
                v *= 1e9
            d[r[ix["Metric Name"]]] = v
    seq = [rows[k] for k in sorted(rows)]
    names = [s["name"] for s in seq]
    # period of the launch sequence = one training step (the optimizer kernel appears once per step)
    marks = [i for i, n in enumerate(names) if n.startswith(("adam_vec", "sgd_vec"))]
    if len(marks) < 2:
        raise SystemExit("less than two optimizer launches in the capture: no full step")
    step = seq[marks[0] + 1: marks[1] + 1]
    T = "gpu__time_duration.sum"
    P = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    fam = collections.OrderedDict()
    per_kernel = collections.OrderedDict()
    for s in step:
        for key, table in ((family(s["name"]), fam), (s["name"], per_kernel)):
            d = table.setdefault(key, dict(launches=0, us=0.0, dram=0.0, l2=0.0, pipe_w=0.0))
            d["launches"] += 1
            d["us"] += s.get(T, 0.0)
            d["dram"] += s.get("dram__bytes_read.sum", 0.0) + s.get("dram__bytes_write.sum", 0.0)
            d["l2"] += s.get("lts__t_bytes.sum", 0.0)
            d["pipe_w"] += s.get(P, 0.0) * s.get(T, 0.0)
    total = sum(d["us"] for d in fam.values())
    print(f"one step = {len(step)} kernel launches, {total:.0f} us of serialised kernel time (cold caches: compare SHARES)")
    print(f"{'kernel':46s} {'n':>4s} {'us':>9s} {'share':>6s} {'DRAM MB':>9s} {'GB/s':>7s} {'L2 MB':>9s} {'tensor pipe %':>13s}")
    for k, d in sorted(per_kernel.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{k:46s} {d['launches']:4d} {d['us']:9.1f} {d['us'] / total:6.3f} {d['dram'] / 1e6:9.1f} "
              f"{d['dram'] / d['us'] / 1e3 if d['us'] else 0:7.0f} {d['l2'] / 1e6:9.1f} {d['pipe_w'] / d['us'] if d['us'] else 0:13.1f}")
    out = {}
    for k, d in fam.items():
        out[k] = dict(dram_bytes_per_step=d["dram"], us_per_step_ncu=d["us"], launches_per_step=d["launches"],
                      share_of_step_ncu=d["us"] / total, pipe_tensor_active_pct=round(d["pipe_w"] / d["us"], 1) if d["us"] else None)
        print(f"FAMILY {k}: {d['launches']} launches, {d['us']:.0f} us ({d['us'] / total:.3f}), DRAM {d['dram'] / 1e6:.0f} MB, "
              f"time-weighted tensor pipe active {d['pipe_w'] / d['us'] if d['us'] else 0:.1f} %")
    out["_source"] = f"{path}: ncu --metrics, one eager step of bench.py --graph off (see profiles/README.md)"
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()

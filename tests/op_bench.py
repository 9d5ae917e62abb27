"""Op-level micro-benchmark through the C ABI (BASELINE.json configs[2]: conv2d / max_pool2d / BN / ReLU sweep).

    python tests/op_bench.py [--ops conv,bn,...] [--shapes r224|r56] [--math tf32|fp32] [--iters 20] [--json out.json]

Each op is timed with CUDA events over `iters` back-to-back launches after 3 warm-ups, on tensors larger than the
126 MB L2 where the layer allows it (else noted "L2-resident"), and reported as algorithmic TFLOP/s or GB/s
(utils/opcost.py) and as a fraction of the measured peaks in MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]

import torch  # noqa: E402

import xshinnosuke_b200 as xs  # noqa: E402
from xshinnosuke_b200._lib import lib  # noqa: E402
from xshinnosuke_b200.core.device import rt  # noqa: E402
from xshinnosuke_b200.utils import opcost  # noqa: E402

# (name, C, H, OC, k, s, p) of tests/resnet ResNet-18 at 224x224 / 56x56 input
LAYERS = {
    "r224": [("stem", 3, 224, 64, 7, 2, 3), ("layer1", 64, 56, 64, 3, 1, 1), ("layer2.0", 64, 56, 128, 3, 2, 1),
             ("layer2", 128, 28, 128, 3, 1, 1), ("layer2.sc", 64, 56, 128, 1, 2, 0), ("layer3.0", 128, 28, 256, 3, 2, 1),
             ("layer3", 256, 14, 256, 3, 1, 1), ("layer3.sc", 128, 28, 256, 1, 2, 0), ("layer4.0", 256, 14, 512, 3, 2, 1),
             ("layer4", 512, 7, 512, 3, 1, 1), ("layer4.sc", 256, 14, 512, 1, 2, 0)],
    "r56": [("stem", 3, 56, 64, 7, 2, 3), ("layer1", 64, 14, 64, 3, 1, 1), ("layer2", 128, 7, 128, 3, 1, 1),
            ("layer3", 256, 4, 256, 3, 1, 1), ("layer4", 512, 2, 512, 3, 1, 1)],
}


def dev(*shape, zero=False):
    t = torch.zeros(*shape, device=rt.device) if zero else torch.randn(*shape, device=rt.device)
    return t


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="conv_fwd,conv_dgrad,conv_wgrad,bn_fwd,bn_bwd,relu_fwd,relu_bwd,add,maxpool_fwd,"
                                     "maxpool_bwd,bnpool,colsum,sgd,adam")
    ap.add_argument("--shapes", default="r224")
    ap.add_argument("--math", default="tf32")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--layers", default="")
    ap.add_argument("--json", default="")
    ap.add_argument("--sweep", default="", choices=["", "small", "full"],
                    help="BASELINE.json configs[2]: conv fwd/dgrad/wgrad and max-pool over N=32-256 x C=3-512 x 32-224 px x "
                         "k=1/3/7 ('small' = one batch size per shape class, 'full' = the whole grid that fits 8 GB per tensor)")
    ap.add_argument("--pers", type=int, default=-1,
                    help="xsb_tc_config(2, v): 0 = per-tile im2col kernels only, 1 = persistent kernel for strided dgrads "
                         "(default), 2 = persistent kernel for every im2col forward / dgrad")
    ap.add_argument("--cpu-ref", action="store_true",
                    help="also time the reference's NumPy algorithm (oracle/, float64 im2col + np.dot + col2im) for conv "
                         "fwd / bwd and BN on a 2-image slice of each layer, on the host cores (context, SURVEY 8d)")
    a = ap.parse_args()
    rt.ensure()
    xs.set_math_mode(a.math)
    if a.pers >= 0:
        lib.xsb_tc_config(2, a.pers)
    peaks = json.load(open(os.path.join(os.path.dirname(HERE), "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(HERE), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    N = a.batch or (128 if a.shapes == "r224" else 32)
    ops = a.ops.split(",")
    st = rt.stream()
    ws, wsn, wss = rt.ws_big.data_ptr(), rt.WS_BIG_FLOATS, rt.ws_small.data_ptr()
    mode = rt.math_mode
    rows = []

    def report(op, layer, name, args, sec, note=""):
        kind, amount = opcost.cost(name, args)
        if kind == "flop":
            ach, unit, peak = amount / sec / 1e12, "TFLOP/s", peaks["bf16_tflops"] / (2.0 if a.math == "tf32" else 1.0)
            pk = "tf32 peak = measured bf16 burst / 2" if a.math == "tf32" else "measured bf16 burst"
        else:
            ach, unit, peak, pk = amount / sec / 1e9, "GB/s", peaks["hbm_gbs"], "measured HBM copy"
        rows.append(dict(op=op, layer=layer, us=sec * 1e6, achieved=ach, unit=unit, frac=ach / peak, peak=peak,
                         peak_kind=pk, note=note))
        print(f"{op:12s} {layer:10s} {sec * 1e6:9.1f} us  {ach:9.1f} {unit:8s} {100 * ach / peak:5.1f}% of {pk} {note}", flush=True)

    def cpu_ref(lname, C, H, OC, k, s_, p_):
        """The oracle (test infrastructure) as the CPU yardstick: per-op GFLOP/s / GB/s of the reference's algorithm."""
        import time
        from oracle import xs_numpy as X
        n = 2
        rs = np.random.RandomState(0)
        xh = rs.randn(n, C, H, H)
        wh = rs.randn(OC, C, k, k) * 0.1
        bh = rs.randn(1, OC)
        t0 = time.perf_counter()
        yh, col = X.conv2d_fwd(xh, wh, bh, (s_, s_), p_)
        t1 = time.perf_counter()
        X.conv2d_bwd(np.ones_like(yh), xh.shape, wh, col, (s_, s_), p_)
        t2 = time.perf_counter()
        fl = 2.0 * yh.size * C * k * k
        X.batch_norm_fwd(yh, np.ones(OC), np.zeros(OC), np.zeros(OC), np.ones(OC), 1, 1e-6, 0.99)
        t3 = time.perf_counter()
        row = dict(op="cpu_ref", layer=lname, images=n, conv_fwd_gflops=fl / (t1 - t0) / 1e9,
                   conv_bwd_gflops=2 * fl / (t2 - t1) / 1e9, bn_fwd_gbs=3 * 8.0 * yh.size / (t3 - t2) / 1e9,
                   threads=os.cpu_count())
        rows.append(row)
        print(f"cpu_ref      {lname:10s} NumPy f64, {n} images: conv fwd {row['conv_fwd_gflops']:7.1f} GFLOP/s, "
              f"conv bwd {row['conv_bwd_gflops']:7.1f} GFLOP/s, BN fwd {row['bn_fwd_gbs']:6.2f} GB/s", flush=True)

    sel = [l + (N,) for l in LAYERS[a.shapes] if not a.layers or l[0] in a.layers.split(",")]
    if a.sweep:
        # configs[2]: N x C x H x k grid (stride 1 for k = 1 / 3, stride 2 for k = 7 as in the stem; pad = k // 2; OC = max(64, C))
        sel = []
        for n_ in ((32, 64, 128, 256) if a.sweep == "full" else (64,)):
            for C_ in (3, 64, 128, 256, 512):
                for H_ in (32, 56, 112, 224):
                    for k_ in (1, 3, 7):
                        OC_ = max(64, C_)
                        s_ = 2 if k_ == 7 else 1
                        OH_ = (H_ - k_ + 2 * (k_ // 2)) // s_ + 1
                        big = 4.0 * n_ * max(H_ * H_ * C_, OH_ * OH_ * OC_)
                        if big > (8e9 if a.sweep == "full" else 2e9) or (C_ == 3 and k_ == 1):
                            continue
                        sel.append((f"N{n_}C{C_}H{H_}k{k_}", C_, H_, OC_, k_, s_, k_ // 2, n_))
        ops = [o for o in ops if o in ("conv_fwd", "conv_dgrad", "conv_wgrad", "maxpool_fwd", "maxpool_bwd")]
    for (lname, C, H, OC, k, s, p, N) in sel:
        OH = (H - k + 2 * p) // s + 1
        x, w, b = dev(N, H, H, C), dev(OC, k, k, C) * 0.1, dev(OC)
        y, dy = dev(N, OH, OH, OC), dev(N, OH, OH, OC)
        dx, dw = dev(N, H, H, C), dev(OC, k, k, C)
        g = (N, H, H, C, OC, k, k, s, s, p)
        if a.cpu_ref:
            cpu_ref(lname, C, H, OC, k, s, p)
        if "conv_fwd" in ops:
            args = (x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr()) + g + (mode, ws, wsn, st)
            report("conv_fwd", lname, "xsb_conv2d_fwd", args, timed(lambda: lib.xsb_conv2d_fwd(*args), a.iters))
        if "conv_fwd_stats" in ops:
            import ctypes
            stats, nparts = torch.zeros(rt.WS_STATS_FLOATS, device=rt.device), ctypes.c_int(0)
            args = (x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr()) + g + (mode, ws, wsn)
            tail = (stats.data_ptr(), rt.WS_STATS_FLOATS, ctypes.byref(nparts), st)
            report("conv_fwd_st", lname, "xsb_conv2d_fwd_stats", args + (st,),
                   timed(lambda: lib.xsb_conv2d_fwd_stats(*args, *tail), a.iters), f"(partials {nparts.value})")
        if "conv_dgrad" in ops and C >= 32:   # (the first layer's input needs no gradient)
            args = (dy.data_ptr(), w.data_ptr(), dx.data_ptr()) + g + (0, mode, ws, wsn, st)
            report("conv_dgrad", lname, "xsb_conv2d_dgrad", args, timed(lambda: lib.xsb_conv2d_dgrad(*args), a.iters))
        if "conv_wgrad" in ops:
            args = (dy.data_ptr(), x.data_ptr(), dw.data_ptr()) + g + (0, mode, ws, wsn, st)
            report("conv_wgrad", lname, "xsb_conv2d_wgrad", args, timed(lambda: lib.xsb_conv2d_wgrad(*args), a.iters))
        # bandwidth ops on the conv OUTPUT tensor [N,OH,OH,OC]
        R, n = N * OH * OH, N * OH * OH * OC
        note = "" if n * 4 > 126e6 else "(L2-resident)"
        gam, bet, rm, rv, sm, si = dev(OC), dev(OC), dev(OC, zero=True), dev(OC) ** 2 + 1, dev(OC), dev(OC)
        if "bn_fwd" in ops:
            args = (y.data_ptr(), gam.data_ptr(), bet.data_ptr(), rm.data_ptr(), rv.data_ptr(), dy.data_ptr(), sm.data_ptr(),
                    si.data_ptr(), R, OC, 1e-6, 0.99, wss, st)
            report("bn_fwd", lname, "xsb_bn_train_fwd", args, timed(lambda: lib.xsb_bn_train_fwd(*args), a.iters), note)
        if "bn_bwd" in ops:
            dxo, dg, db = dev(N, OH, OH, OC), dev(OC), dev(OC)
            args = (dy.data_ptr(), y.data_ptr(), gam.data_ptr(), sm.data_ptr(), si.data_ptr(), dxo.data_ptr(), dg.data_ptr(),
                    db.data_ptr(), R, OC, 0, 0, 0, None, wss, st)
            report("bn_bwd", lname, "xsb_bn_bwd", args, timed(lambda: lib.xsb_bn_bwd(*args), a.iters), note)
        if "colsum" in ops:
            args = (dy.data_ptr(), b.data_ptr(), R, OC, 0, wss, st)
            report("colsum", lname, "xsb_colsum", args, timed(lambda: lib.xsb_colsum(*args), a.iters), note)
        mask = torch.zeros((n + 7) // 8, dtype=torch.uint8, device=rt.device)
        if "relu_fwd" in ops:
            args = (y.data_ptr(), y.data_ptr(), mask.data_ptr(), n, st)
            report("relu_fwd", lname, "xsb_relu_fwd", args, timed(lambda: lib.xsb_relu_fwd(*args), a.iters), note)
        if "relu_bwd" in ops:
            args = (dy.data_ptr(), mask.data_ptr(), dy.data_ptr(), n, 0, st)
            report("relu_bwd", lname, "xsb_relu_bwd_mask", args, timed(lambda: lib.xsb_relu_bwd_mask(*args), a.iters), note)
        if "add" in ops:
            args = (y.data_ptr(), dy.data_ptr(), y.data_ptr(), n, st)
            report("add", lname, "xsb_add", args, timed(lambda: lib.xsb_add(*args), a.iters), note)
        if a.sweep and k == 3 and C >= 64 and ("maxpool_fwd" in ops or "maxpool_bwd" in ops):
            for pk_, pp_ in ((2, 0), (3, 1)):          # max-pool k in {2, 3}, stride 2, pad in {0, 1} on the INPUT tensor
                PH = (H + 2 * pp_ - pk_) // 2 + 1
                py, am = dev(N, PH, PH, C), torch.zeros(N * PH * PH * C, dtype=torch.uint8, device=rt.device)
                xr = torch.relu(x)
                args = (xr.data_ptr(), py.data_ptr(), am.data_ptr(), N, H, H, C, pk_, 2, 2, pp_, st)
                if "maxpool_fwd" in ops:
                    report("maxpool_fwd", f"{lname}p{pk_}", "xsb_maxpool2d_fwd", args, timed(lambda: lib.xsb_maxpool2d_fwd(*args), a.iters))
                lib.xsb_maxpool2d_fwd(*args)
                args2 = (py.data_ptr(), am.data_ptr(), dx.data_ptr(), N, H, H, C, pk_, 2, 2, pp_, 0, st)
                if "maxpool_bwd" in ops:
                    report("maxpool_bwd", f"{lname}p{pk_}", "xsb_maxpool2d_bwd", args2, timed(lambda: lib.xsb_maxpool2d_bwd(*args2), a.iters))
                del py, am, xr
        if not a.sweep and lname == "stem" and ("maxpool_fwd" in ops or "maxpool_bwd" in ops or "bnpool" in ops):
            PH = (OH + 2 - 3) // 2 + 1
            py, am = dev(N, PH, PH, OC), torch.zeros(N * PH * PH * OC, dtype=torch.uint8, device=rt.device)
            yr = torch.relu(y)
            args = (yr.data_ptr(), py.data_ptr(), am.data_ptr(), N, OH, OH, OC, 3, 2, 2, 1, st)
            if "maxpool_fwd" in ops:
                report("maxpool_fwd", lname, "xsb_maxpool2d_fwd", args, timed(lambda: lib.xsb_maxpool2d_fwd(*args), a.iters))
            lib.xsb_maxpool2d_fwd(*args)
            args2 = (py.data_ptr(), am.data_ptr(), dy.data_ptr(), N, OH, OH, OC, 3, 2, 2, 1, 0, st)
            if "maxpool_bwd" in ops:
                report("maxpool_bwd", lname, "xsb_maxpool2d_bwd", args2, timed(lambda: lib.xsb_maxpool2d_bwd(*args2), a.iters))
            if "bnpool" in ops:      # the stem's BN -> ReLU -> pool as fused kernels (forward: one launch, backward: two)
                del yr
                args3 = (y.data_ptr(), sm.data_ptr(), si.data_ptr(), gam.data_ptr(), bet.data_ptr(), py.data_ptr(), am.data_ptr(),
                         mask.data_ptr(), N, OH, OH, OC, 3, 2, 2, 1, st)
                report("bnpool_fwd", lname, "xsb_bn_relu_maxpool_fwd", args3,
                       timed(lambda: lib.xsb_bn_relu_maxpool_fwd(*args3), a.iters), "(replaces bn_apply+relu and maxpool_fwd)")
                dxo, dg, db = dev(N, OH, OH, OC), dev(OC), dev(OC)
                args4 = (py.data_ptr(), am.data_ptr(), y.data_ptr(), gam.data_ptr(), sm.data_ptr(), si.data_ptr(), dxo.data_ptr(),
                         dg.data_ptr(), db.data_ptr(), N, OH, OH, OC, 0, 0, 0, mask.data_ptr(), wss, None, 0, st)
                report("bnpool_bwd", lname, "xsb_bn_bwd_pool_dxsum", args4,
                       timed(lambda: lib.xsb_bn_bwd_pool_dxsum(*args4), a.iters), "(replaces maxpool_bwd and bn_bwd)")
        del x, w, y, dy, dx, dw
        torch.cuda.empty_cache()
    P = 11232612 // 64 * 64 + 64 * 82
    fp, fg, fv, fm = dev(P), dev(P), dev(P, zero=True), dev(P, zero=True)
    stp = torch.ones(1, dtype=torch.int32, device=rt.device)
    if "sgd" in ops:
        args = (fp.data_ptr(), fg.data_ptr(), P, 0.1, 1.0, 0.0, st)
        report("sgd", "flat", "xsb_sgd_step", args, timed(lambda: lib.xsb_sgd_step(*args), a.iters), "(90 MB: L2-resident)")
    if "adam" in ops:
        args = (fp.data_ptr(), fg.data_ptr(), fv.data_ptr(), fm.data_ptr(), P, 1e-3, 0.9, 0.999, 1e-8, 1.0, 0.0, stp.data_ptr(), st)
        report("adam", "flat", "xsb_adam_step", args, timed(lambda: lib.xsb_adam_step(*args), a.iters))
    if a.json:
        json.dump(dict(shapes=a.shapes, batch=N, math=a.math, iters=a.iters, rows=rows), open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()

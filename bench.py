#!/usr/bin/env python
"""bench.py -- training throughput of the xs hot path on N B200s, one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W [--workload r224|r56|cnn|mlp] [--math tf32|fp32|fp32x3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's NumPy algorithm (oracle port) on the host cores

A "step" = zero_grad, forward, cross-entropy, backward, (bucketed all-reduce), fused optimizer step on one batch of
synthetic data through the public xs API of xshinnosuke_b200 (the loop of the reference's
tests/resnet/resnet18_dg.py:92-100, resp. the body of Model.fit, xs/nn/models.py:100-106). `value` keeps the batch
resident in HBM; `e2e` uploads every step's batch from pinned host memory and reads the loss back.

Workloads (BASELINE.json configs): r224 = configs[4] per GPU (default, the configuration the headline metric is quoted
on), r56 = configs[3], cnn = configs[0], mlp = configs[1]; configs[2] (op sweep) is tests/op_bench.py --sweep.
See DESIGN.md "Measurement" for what every key means and how the kernel table is timed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[3] (3x32x32 does not run on the reference: min input 33, SURVEY.md 8c -> its own 56x56)
    "r56": dict(kind="resnet", batch=32, hw=56, optimizer="sgd", lr=0.1, classes=100, metric="resnet18_train_images_per_sec",
                name="ResNet-18 (tests/resnet), batch 32/GPU, synthetic 3x56x56, SGD lr 0.1, dynamic graph"),
    # BASELINE.json configs[4]
    "r224": dict(kind="resnet", batch=128, hw=224, optimizer="adam", lr=1e-3, classes=100, metric="resnet18_train_images_per_sec",
                 name="ResNet-18 (tests/resnet), batch 128/GPU, synthetic 3x224x224, Adam lr 1e-3, dynamic graph"),
    # BASELINE.json configs[0] in the form that runs on the reference
    "cnn": dict(kind="cnn", batch=256, optimizer="sgd", lr=0.01, classes=10, metric="readme_cnn_train_images_per_sec",
                name="README functional CNN Input(1,28,28)-Conv2D(8,2x2)-relu-MaxPooling2D(2)-Flatten-Dense(10), batch 256/GPU, "
                     "SGD lr 0.01, static graph (Model)"),
    # BASELINE.json configs[1]
    "mlp": dict(kind="mlp", batch=128, optimizer="sgd", lr=0.05, classes=10, metric="mlp_train_images_per_sec",
                name="Sequential MLP Dense(784->500, relu)-Dense(10), batch 128/GPU, SGD lr 0.05, static graph (Sequential)"),
}
README_GPU_IMG_S = 265.0   # XS dynamic graph on CuPy, unspecified GPU (README.md:37-38)
TENSOR_ENTRIES = ("xsb_conv2d_fwd", "xsb_conv2d_fwd_stats", "xsb_conv2d_fwd_acc", "xsb_conv2d_dgrad", "xsb_conv2d_wgrad", "xsb_gemm")
CONV_ENTRIES = TENSOR_ENTRIES[:5]
PARITY_TOL_NET = 5e-3      # whole network in tf32 (20 tf32 convolutions deep); per-op bar is 2e-3 (tests/test_parity_gpu.py)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        out = dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d["bf16_tflops_sustained"],
                   source="MEASURED_PEAKS.json")
    else:
        out = dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback (B200_PROFILING.md)")
    # dense TF32 measured the same way (tools/measure_tf32_peak.py, cuBLAS 8192^3): the conv kernels run kind::tf32
    t32 = os.path.join(ROOT, "profiles", "tf32_peak.json")
    if os.path.exists(t32):
        t = json.load(open(t32))
        out.update(tf32_burst=t["tf32_tflops"], tf32_sustained=t["tf32_tflops_sustained"])
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.12)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    power_w_max=max(power) if power else None, reasons=sorted(reasons), samples=len(sm))


def signature(name, a):
    """Kernel + shape key (conv geometry / GEMM dims / element count)."""
    try:
        if name in ("xsb_conv2d_fwd", "xsb_conv2d_fwd_stats", "xsb_conv2d_fwd_acc"):
            N, H, W, C, OC, KH, KW, sh, sw, pad = a[4:14]
        elif name in ("xsb_conv2d_dgrad", "xsb_conv2d_wgrad"):
            N, H, W, C, OC, KH, KW, sh, sw, pad = a[3:13]
        elif name == "xsb_gemm":
            return f"{name}[M{a[4]} N{a[5]} K{a[6]}]"
        elif name in ("xsb_bn_stats",):
            return f"{name}[R{a[5]} C{a[6]}]"
        elif name in ("xsb_bn_apply",):
            return f"{name}[R{a[6]} C{a[7]} relu{a[8]}]"
        elif name in ("xsb_bn_bwd", "xsb_bn_train_fwd", "xsb_bn_bwd_dxsum", "xsb_bn_bwd_dxsum_side"):
            return f"{name}[R{a[8]} C{a[9]}]"
        elif name == "xsb_colsum":
            return f"{name}[R{a[2]} C{a[3]}]"
        else:
            return name
        return f"{name}[N{N} H{H} W{W} C{C} OC{OC} k{KH}x{KW} s{sh} p{pad}]"
    except Exception:
        return name


# ---------------------------------------------------------------------------------------------- CPU arms (oracle)
def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every host core."""
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=os.cpu_count())
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count()


def synthetic_host_batch(wl, n, seed):
    """Host batch of the workload's shape: images rand [0,1) fp32, labels randint."""
    rs = np.random.RandomState(seed)
    if wl["kind"] == "resnet":
        x = rs.rand(n, 3, wl["hw"], wl["hw"])
    elif wl["kind"] == "cnn":
        x = rs.rand(n, 1, 28, 28)
    else:
        x = rs.rand(n, 784)
    return x.astype(np.float32), rs.randint(0, wl["classes"], (n,)).astype(np.int64)


def oracle_net(wl, weight_seed):
    """The float64 NumPy restatement of the reference path for this workload (oracle/ -- test infrastructure; here it is
    the thing TIMED on the host cores and the checker of `parity_checked`, never part of the product path)."""
    np.random.seed(weight_seed)
    if wl["kind"] == "resnet":
        from oracle.resnet18 import ResNet18 as OracleNet
        net, state = OracleNet(num_classes=wl["classes"]), {}
        if wl["optimizer"] == "sgd":
            return net, (lambda x, y: net.train_step(x, y, wl["lr"]))
        return net, (lambda x, y: net.train_step(x, y, wl["lr"], "adam", state))
    from oracle.small_nets import ReadmeCNN, ReadmeMLP
    net = ReadmeCNN() if wl["kind"] == "cnn" else ReadmeMLP()
    return net, (lambda x, y: net.train_step(x, y, wl["lr"]))


def cpu_sample_size(wl):
    # bounded sample: a step on the full 224x224 batch is ~1.4 TFLOP of float64 -> a slice of the batch instead
    return 2 if (wl["kind"] == "resnet" and wl["hw"] > 64) else wl["batch"]


def time_oracle(wl, budget_s, max_steps, min_steps=2):
    cores = _use_all_host_threads()
    n = cpu_sample_size(wl)
    x, y = synthetic_host_batch(wl, n, seed=0)
    x = x.astype(np.float64)
    _, step = oracle_net(wl, weight_seed=0)
    step(x, y)
    t0, k = time.perf_counter(), 0
    while k < min_steps or (time.perf_counter() - t0 < budget_s and k < max_steps):
        step(x, y)
        k += 1
    dt = (time.perf_counter() - t0) / k
    shape = "x".join(str(s) for s in x.shape[1:])
    return dict(value=n / dt, unit="images/s", cores=cores, kind="port",
                sample=f"{k} steps of {n} samples ({shape}) through oracle/ (float64 NumPy restatement of the reference "
                       f"path: im2col + np.dot + col2im), {dt * 1e3:.1f} ms/step"), dt, k, n


def run_reference(args, wl):
    """The reference's own algorithm (oracle/ port of xs/nn/functional.py, xs/nn/grad_fn.py, xs/optim; the reference itself
    is pure Python and cannot travel to the GPU box) on the host cores, every BLAS thread it can use. Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cpu, dt, k, n = time_oracle(wl, budget_s=60.0, max_steps=max(1, min(args.steps, 200)), min_steps=min(2, max(1, args.steps)))
    out = dict(metric=wl["metric"], value=cpu["value"], unit="images/s", n_gpus=args.gpus, steps=k,
               warmup=1, ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f64", data="synthetic", impl="reference",
               config=dict(workload=wl["name"], sample_batch=n), cpu_baseline=cpu,
               e2e=dict(value=cpu["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(out)


# ---------------------------------------------------------------------------------------------- our arm
def build_workload(wl, xs, nn, x_host, y_host):
    """-> (eager_step, x_dev, y_dev, net_parameters, optimizer, model) through the public xs API."""
    x_dev, y_dev = xs.tensor(x_host), xs.tensor(y_host)
    if wl["kind"] == "resnet":
        from xshinnosuke_b200.recipes import ResNet18
        net = ResNet18(num_classes=wl["classes"])
        opt = xs.optim.SGD(net.parameters(), lr=wl["lr"]) if wl["optimizer"] == "sgd" else \
            xs.optim.Adam(net.parameters(), lr=wl["lr"])
        crit = nn.CrossEntropyLoss()

        def step():
            opt.zero_grad()
            loss = crit(net(x_dev), y_dev)
            loss.backward()
            opt.step()
            return loss
        return step, x_dev, y_dev, net, opt
    from xshinnosuke_b200.recipes import readme_cnn, readme_mlp
    model = readme_cnn(lr=wl["lr"]) if wl["kind"] == "cnn" else readme_mlp(lr=wl["lr"])
    model.init_graph_out_tensor(x_dev, y_dev)

    def step():          # the body of Model.fit's batch loop (xs/nn/models.py:100-106) on the static graph
        model.train()
        model.optimizer.zero_grad()
        pred = model.forward(x_dev)
        loss = model.loss.forward(pred, y_dev)
        model.backward(loss)
        model.optimizer.step()
        return loss
    return step, x_dev, y_dev, model, model.optimizer


def parity_check(wl, args, xs, nn):
    """Step 0 of the GPU path (the bench's math mode) against the float64 oracle on the SAME bounded sample the CPU arm
    times, from identical seeded weights: loss, logits, the LAST layer's weight gradient (one product away from the loss)
    and the FIRST layer's (through the whole backward pass). The first-layer figure is informational: a deep, randomly
    initialised ResNet amplifies any rounding (ReLU-mask and arg-max flips) -- the float64 oracle itself moves by
    `oracle_tf32_model_*` when its GEMM operands are cut to tf32 (oracle.xs_numpy.TF32_OPERANDS), and the pure-fp32 FFMA
    path shows 1e-3..1e-2 there as well (tools/parity_probe.py)."""
    import xshinnosuke_b200.core.state as G
    from oracle import xs_numpy as X
    n = cpu_sample_size(wl)
    x, y = synthetic_host_batch(wl, n, seed=0)
    x64 = x.astype(np.float64)

    def oracle_step0(tf32_operands):
        X.TF32_OPERANDS = tf32_operands
        try:
            onet, _ = oracle_net(wl, weight_seed=0)
            if wl["kind"] == "resnet":
                logits = onet.forward(x64)
                loss, probs = X.cross_entropy_fwd(logits, y)
                grads = onet.backward(X.cross_entropy_bwd(probs, y))
                return dict(loss=float(loss[0]), logits=logits, dw_first=grads[0], dw_last=grads[-2])
            w0 = [p.copy() for p in onet.params]
            lossv = onet.train_step(x64, y, wl["lr"])            # SGD: dW = (W0 - W1) / lr
            return dict(loss=lossv, logits=None, dw_first=(w0[0] - onet.params[0]) / wl["lr"],
                        dw_last=(w0[-2] - onet.params[-2]) / wl["lr"])
        finally:
            X.TF32_OPERANDS = False
    ref = oracle_step0(False)
    G.INPUTS = G.OUTPUTS = G.GRAPH = None
    np.random.seed(0)
    tx, ty = xs.tensor(x), xs.tensor(y)
    got_logits = None
    if wl["kind"] == "resnet":
        from xshinnosuke_b200.recipes import ResNet18
        net = ResNet18(num_classes=wl["classes"])
        opt = xs.optim.SGD(net.parameters(), lr=wl["lr"])
        opt.zero_grad()
        out = net(tx)
        got_logits = out.eval
        loss = nn.CrossEntropyLoss()(out, ty)
        got_loss = loss.item()
        loss.backward()
        params = sorted(net.parameters(), key=lambda p: p.ticket)
    else:
        from xshinnosuke_b200.recipes import readme_cnn, readme_mlp
        model = readme_cnn(lr=wl["lr"]) if wl["kind"] == "cnn" else readme_mlp(lr=wl["lr"])
        model.init_graph_out_tensor(tx, ty)
        model.train()
        model.optimizer.zero_grad()
        pred = model.forward(tx)
        loss = model.loss.forward(pred, ty)
        got_loss = loss.item()
        model.backward(loss)
        params = sorted(model.parameters(), key=lambda p: p.ticket)
    got_first, got_last = params[0].grad.eval, params[-2].grad.eval
    G.INPUTS = G.OUTPUTS = G.GRAPH = None
    errs = dict(loss_rel_err=abs(got_loss - ref["loss"]) / abs(ref["loss"]),
                last_layer_dw_rel_err=float(X.rel_err(got_last.reshape(ref["dw_last"].shape), ref["dw_last"])))
    if ref["logits"] is not None:
        errs["logits_rel_err"] = float(X.rel_err(got_logits, ref["logits"]))
    tol = 1e-5 if args.math in ("fp32", "fp32x3") else PARITY_TOL_NET
    # fp32 modes through a whole network: fp32 storage rounding of every activation on the way (the per-op bar is 1e-5)
    net_tol = max(tol, 1e-4) if wl["kind"] == "resnet" else max(tol, 2e-5)
    tols = {k: net_tol for k in errs}
    info = dict(first_layer_dw_rel_err=float(X.rel_err(got_first.reshape(ref["dw_first"].shape), ref["dw_first"])))
    if args.math == "tf32" and wl["kind"] == "resnet":
        # how far the float64 oracle ITSELF moves when only its GEMM operands are cut to tf32: the floor of any tf32
        # implementation through 18 layers. A quantity passes within max(5e-3, 2 x that floor) -- the per-op 2e-3 bar is
        # what tests/test_parity_gpu.py enforces; this check is there to catch a broken kernel at the benched shape
        model_ref = oracle_step0(True)
        floor = dict(loss_rel_err=abs(model_ref["loss"] - ref["loss"]) / abs(ref["loss"]),
                     logits_rel_err=float(X.rel_err(model_ref["logits"], ref["logits"])),
                     last_layer_dw_rel_err=float(X.rel_err(model_ref["dw_last"], ref["dw_last"])))
        for k, v in floor.items():
            tols[k] = max(net_tol, 2.0 * v)
        info["oracle_tf32_model_first_layer_dw_dev"] = float(X.rel_err(model_ref["dw_first"], ref["dw_first"]))
        info["oracle_tf32_model_logits_dev"] = floor["logits_rel_err"]
        info["oracle_tf32_model_last_layer_dw_dev"] = floor["last_layer_dw_rel_err"]
        info["first_layer_dw_vs_tf32_model"] = float(X.rel_err(got_first.reshape(ref["dw_first"].shape), model_ref["dw_first"]))
    passed = bool(all(v <= tols[k] for k, v in errs.items()))
    tol_by = {k: float(t) for k, t in tols.items()}
    errs.update(info)
    errs.update(sample=f"{n} samples, step 0, identical seeded weights, float64 oracle", math_mode=args.math, tol=net_tol,
                tol_by_quantity=tol_by,
                tol_applies_to="loss, logits, last_layer_dw (first_layer_dw is informational: chaotic amplification)",
                metric="max|a-b|/max|b|", passed=passed)
    return errs


KERNEL_LIST = []     # (C-ABI call signature, device kernel name, eager us, graph-replay us) of one step, launch order


def kernel_timeline(eager_step, graphed, lib, torch, n_replays=3):
    """Per C-ABI call kernel time measured INSIDE the CUDA-graph replay.

    (1) one eager step under the profiler with every C-ABI call wrapped in a record_function range: CUPTI attributes each
    kernel to the call that launched it; (2) `n_replays` graph replays under the profiler: the replay runs the same kernels in
    the same order, so kernel i of a replay is kernel i of the eager list and takes ITS replay duration. Falls back to the
    eager kernel durations if the two lists do not line up. -> (records [(entry, args, seconds)], how)."""
    from torch.profiler import ProfilerActivity, profile, record_function
    from torch.autograd import DeviceType
    names = [n for n in lib.protos if n.startswith("xsb_") and lib.protos[n][1] and
             n not in ("xsb_init", "xsb_stream_sync", "xsb_memcpy_h2d", "xsb_memcpy_d2h", "xsb_bn_workspace_floats",
                       "xsb_tc_stats", "xsb_tc_config", "xsb_tc_conv_supported")]
    originals, calls = {}, []
    for n in names:
        fn = getattr(lib, n)
        originals[n] = fn

        def wrapped(*a, _fn=fn, _n=n):
            idx = len(calls)
            calls.append((_n, a))
            with record_function(f"xsbcall::{idx}"):
                return _fn(*a)
        setattr(lib, n, wrapped)
    try:
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            eager_step()
            torch.cuda.synchronize()
    finally:
        for n, fn in originals.items():
            setattr(lib, n, fn)
    # attribution: kernel --(correlation id)--> its cudaLaunchKernel runtime event --(launch time inside the range)--> C-ABI call
    import bisect
    evs = list(prof.events())
    ranges = sorted((ev.time_range.start, ev.time_range.end, int(ev.name.split("::")[1])) for ev in evs
                    if ev.name.startswith("xsbcall::") and ev.device_type == DeviceType.CPU)   # (not the GPU-side twin)
    starts = [r[0] for r in ranges]
    launch_t = {ev.id: ev.time_range.start for ev in evs
                if ev.device_type == DeviceType.CPU and ev.name.startswith(("cudaLaunch", "cuLaunch"))}
    eager_kernels = []          # (call index, kernel name, eager duration us) in launch order
    dev = [ev for ev in evs if ev.device_type == DeviceType.CUDA and
           not ev.name.lower().startswith(("memcpy", "memset", "xsbcall::"))]      # (ranges have a GPU-side twin)
    dev.sort(key=lambda ev: ev.time_range.start)
    for ev in dev:
        t = launch_t.get(ev.id)
        if t is None:
            continue
        j = bisect.bisect_right(starts, t) - 1
        if j >= 0 and t <= ranges[j][1]:
            eager_kernels.append((ranges[j][2], ev.name, ev.time_range.elapsed_us()))
    how = "eager kernel durations (CUPTI), attributed to C-ABI calls"
    durs = [d for _, _, d in eager_kernels]
    if graphed is not None and eager_kernels:
        with profile(activities=[ProfilerActivity.CUDA]) as prof2:
            for _ in range(n_replays):
                graphed()
            torch.cuda.synchronize()
        ks = [ev for ev in prof2.events() if ev.device_type == DeviceType.CUDA and
              not ev.name.lower().startswith(("memcpy", "memset"))]
        ks.sort(key=lambda ev: ev.time_range.start)
        L = len(eager_kernels)
        # the graph holds the step's kernels only; library fills (at::FillFunctor) launched outside C-ABI calls are
        # in the replay as well: align by name, skipping replay kernels the eager attribution does not know
        if len(ks) >= n_replays * L:
            per = len(ks) // n_replays
            if per * n_replays == len(ks):
                acc, ok = [0.0] * L, True
                for r in range(n_replays):
                    seq, i = ks[r * per:(r + 1) * per], 0
                    for ev in seq:
                        if i < L and ev.name == eager_kernels[i][1]:
                            acc[i] += ev.time_range.elapsed_us()
                            i += 1
                    ok = ok and i == L
                if ok:
                    durs = [a / n_replays for a in acc]
                    how = f"CUDA-graph replay kernel durations (CUPTI, mean of {n_replays} replays), attributed through the eager launch order"
    how += f" [{len(eager_kernels)} kernels of {len(calls)} C-ABI calls; {len(dev)} device kernels, {len(launch_t)} launch records seen]"
    by_call = {}
    for (idx, _, _), d in zip(eager_kernels, durs):
        by_call[idx] = by_call.get(idx, 0.0) + d
    KERNEL_LIST[:] = [(signature(calls[idx][0], calls[idx][1]), kname, ed, d) for (idx, kname, ed), d in zip(eager_kernels, durs)]
    return [(calls[i][0], calls[i][1], by_call.get(i, 0.0) * 1e-6) for i in range(len(calls))], how


def run_ours(args, wl):
    import torch
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.core.device import rt
    from xshinnosuke_b200.utils import opcost

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    rt.ensure()
    xs.set_math_mode(args.math)
    parity = None
    dp = None
    B = wl["batch"]
    x_host, y_host = synthetic_host_batch(wl, B, seed=1000 + rank)
    np.random.seed(0)   # identical initial weights on every rank
    eager_step, x_dev, y_dev, net, opt = build_workload(wl, xs, nn, x_host, y_host)
    if world > 1:
        from xshinnosuke_b200.parallel import DataParallel
        dp = DataParallel(net, opt, bucket_mb=args.bucket_mb)
    x_pin = torch.from_numpy(x_host).pin_memory()
    y_pin = torch.from_numpy(y_host).pin_memory()

    def barrier():
        if dp is not None:
            dp.dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dp is not None:
            t = torch.tensor([ms], device=rt.device)
            dp.dist.all_reduce(t, op=dp.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    use_graph = args.graph in ("on", "auto")   # NCCL bucket all-reduces are captured together with the kernels
    step_resident, graphed = eager_step, None
    xh, yh = x_pin.numpy(), y_pin.numpy()
    if use_graph:
        # the whole step replayed as one CUDA graph (same kernels, no Python between launches)
        graphed = xs.GraphedStep(eager_step, warmup=max(args.warmup, 3), inputs=(x_dev, y_dev))
        step_resident = graphed
        graphed.prefetch(xh, yh)

        def step_e2e():  # one H2D upload of a pinned host batch and one D2H loss read per step
            loss = graphed(xh, yh)       # staged batch -> captured inputs (device transpose), replay
            graphed.prefetch(xh, yh)     # the NEXT step's upload runs on a side stream under this replay
            return loss.item()
    else:
        def step_e2e():                  # the call a user makes: host batch in, loss value out
            x_dev.eval = xh
            y_dev.eval = yh
            return eager_step().item()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    l0 = rt.launches
    ms = timed(step_resident, args.steps)
    launches = rt.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- kernel table: time of every C-ABI call's kernels inside the graph replay (CUPTI), rank 0 of a 1-GPU run
    records, how = [], None
    if not args.no_kernel_table:
        lib.load()
        try:
            if world == 1:
                records, how = kernel_timeline(eager_step, graphed, lib, torch)
            elif rank == 0:          # N > 1: one eager step (it contains collectives: every rank takes part), no replay alignment
                records, how = kernel_timeline(eager_step, None, lib, torch)
            else:
                eager_step()
                torch.cuda.synchronize()
        except Exception as e:       # profiler unavailable: the table is omitted, the headline numbers stand
            how = f"unavailable ({type(e).__name__}: {e})"
    if rank != 0:
        return
    # parity at the benched configuration -- AFTER the timed regions: the float64 oracle leaves every BLAS thread of the host
    # spinning, which starves the thread that enqueues the graph replays of a launch-bound workload (r56: -6 %)
    if world == 1 and not args.no_parity:
        import xshinnosuke_b200.core.state as G
        G.INPUTS = G.OUTPUTS = G.GRAPH = None
        parity = parity_check(wl, args, xs, nn)

    pk = peaks()
    prof, by_shape = {}, {}
    for n, a, sec in records:
        kind, amount = opcost.cost(n, a)
        for key, table in ((n, prof), (signature(n, a), by_shape)):
            d = table.setdefault(key, dict(s=0.0, calls=0, amount=0.0, kind=kind, name=n))
            d["s"] += sec
            d["calls"] += 1
            d["amount"] += amount
    total_s = sum(d["s"] for d in prof.values())
    ncu = {}
    npath = os.path.join(ROOT, "profiles", "ncu_metrics.json")
    if os.path.exists(npath):
        ncu = json.load(open(npath))

    def rate(d):
        return d["amount"] / d["s"] / (1e12 if d["kind"] == "flop" else 1e9) if d["s"] > 0 else 0.0

    def tensor_roofline(entries, label):
        ds = [prof[e] for e in entries if e in prof]
        if not ds or sum(d["s"] for d in ds) <= 0:
            return None
        sec, flop = sum(d["s"] for d in ds), sum(d["amount"] for d in ds)
        achieved = flop / sec / 1e12
        # the step is seconds-long back-to-back work for the part: the sustained bf16 figure is the denominator;
        # kind::tf32 runs at half the bf16 rate -> frac_of_tf32_peak is the fraction of what these kernels can reach
        peak = pk["tc_sustained"]
        tf32 = args.math in ("tf32", "fp32x3")
        r = dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                 traffic=ncu.get(label, {}).get("dram_bytes_per_step"), kernel=label,
                 us_per_step=sec * 1e6, share_of_kernel_time=sec / total_s if total_s else None,
                 algorithmic_flop_per_step=flop, peak_source=pk["source"] + " (bf16 sustained)",
                 frac_of_tf32_peak=achieved / (peak / 2.0) if tf32 else None,
                 frac_of_bf16_burst=achieved / pk["tc_burst"],
                 tf32_cublas_measured=pk.get("tf32_sustained"),
                 frac_of_tf32_cublas_measured=achieved / pk["tf32_sustained"] if (tf32 and pk.get("tf32_sustained")) else None,
                 pipe_tensor_active_pct=ncu.get(label, {}).get("pipe_tensor_active_pct"),
                 by_entry={e: dict(tflops=round(rate(prof[e]), 1), us_per_step=round(prof[e]["s"] * 1e6, 1),
                                   calls=prof[e]["calls"]) for e in entries if e in prof},
                 note="achieved = algorithmic FLOPs of every convolution of the step / their summed kernel time in the graph "
                      "replay; in fp32x3 mode each product is three tf32 MMAs (algorithmic FLOPs counted once)")
        return r

    def hbm_roofline():
        hb = {n: d for n, d in prof.items() if d["kind"] == "byte" and d["s"] > 0}
        if not hb:
            return None
        # the BatchNorm backward is one kernel family behind several entry points (plain, + bias-gradient sums, + the
        # identity shortcut's masked copy, + the pool-gathering form): judged as ONE entry
        fam = [n for n in hb if n.startswith("xsb_bn_bwd")]
        if len(fam) > 1:
            merged = dict(s=sum(hb[n]["s"] for n in fam), calls=sum(hb[n]["calls"] for n in fam),
                          amount=sum(hb[n]["amount"] for n in fam), kind="byte", name="xsb_bn_bwd_dxsum")
            for n in fam:
                del hb[n]
            hb["xsb_bn_bwd_dxsum"] = merged
        n, d = max(hb.items(), key=lambda kv: kv[1]["s"])
        achieved = rate(d)
        label = n + (" (all shapes of the step)" if len(fam) <= 1 or n != "xsb_bn_bwd_dxsum" else
                     " family (" + ", ".join(sorted(fam)) + "; all shapes of the step)")
        return dict(bound="hbm", achieved=achieved, peak=pk["hbm"], unit="GB/s", frac=achieved / pk["hbm"],
                    traffic=ncu.get(n, {}).get("dram_bytes_per_step"), kernel=label,
                    us_per_step=d["s"] * 1e6, calls_per_step=d["calls"], share_of_kernel_time=d["s"] / total_s,
                    algorithmic_bytes_per_step=d["amount"], peak_source=pk["source"] + " (HBM copy)")

    roof_t = tensor_roofline([e for e in CONV_ENTRIES], "conv family (xsb_conv2d_fwd + dgrad + wgrad)")
    if roof_t is None:
        roof_t = tensor_roofline(["xsb_gemm"], "Dense products (xsb_gemm)")
    roof_h = hbm_roofline()
    t_share = (roof_t or {}).get("share_of_kernel_time") or 0.0
    h_share = sum(d["s"] for d in prof.values() if d["kind"] == "byte") / total_s if total_s else 0.0
    roofline, other = (roof_t, roof_h) if t_share >= 0.3 or roof_h is None else (roof_h, roof_t)
    per_kernel = {n: dict(share=round(d["s"] / total_s, 4), us_per_step=round(d["s"] * 1e6, 1), calls_per_step=d["calls"],
                          achieved=round(rate(d), 2), unit="TFLOP/s" if d["kind"] == "flop" else "GB/s")
                  for n, d in sorted(prof.items(), key=lambda kv: -kv[1]["s"])} if total_s else {}
    top_shapes = {k: dict(us_per_step=round(d["s"] * 1e6, 1), calls=d["calls"], achieved=round(rate(d), 1),
                          unit="TFLOP/s" if d["kind"] == "flop" else "GB/s")
                  for k, d in sorted(by_shape.items(), key=lambda kv: -kv[1]["s"])[:24]} if total_s else {}

    nparams = sum(p.numel() for p in net.parameters())
    value = B * world * args.steps / (ms * 1e-3)
    e2e_val = B * world * args.steps / (ms_e2e * 1e-3)
    cpu = time_oracle(wl, budget_s=20.0, max_steps=200)[0] if (world == 1 and not args.no_cpu_baseline) else None
    shape = "x".join(str(s) for s in x_host.shape[1:])
    out = dict(metric=wl["metric"], value=value, unit="images/s", n_gpus=world, steps=args.steps,
               warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
               vs_baseline=value / README_GPU_IMG_S if args.workload == "r56" else None,
               dtype={"tf32": "tf32", "fp32": "f32", "fp32x3": "f32 (3xTF32 split on tcgen05)"}[args.math], data="synthetic",
               config=dict(workload=wl["name"], global_batch=B * world, per_gpu_batch=B, sample_shape=shape,
                           optimizer=wl["optimizer"], parallelism=f"dp{world}", params=nparams,
                           math_mode=args.math, cuda_graph=bool(use_graph),
                           l2="no flush: per-step working set (activations + parameters/gradients/moments) exceeds the 126 MB L2"
                           if wl["kind"] == "resnet" else
                           "no flush: the whole step is L2-resident by nature (launch-latency bound workload); the kernels "
                           "of one step never re-read a buffer they did not just produce",
                           vs_baseline_ref="README.md:37-38 XS dynamic graph on CuPy, 265 img/s, unspecified GPU"),
               clocks=clocks, gpu_launches=launches,
               gpu_launches_unit="C-ABI entry calls of libxsb200.so in the timed region (a lower bound on kernels: a BatchNorm "
                                 "backward is 2 launches, a convolution may add a filter re-layout)",
               e2e=dict(value=e2e_val, unit="images/s", h2d_bytes_per_step=int(x_host.nbytes + y_host.nbytes),
                        d2h_bytes_per_step=4, ms_per_step=ms_e2e / args.steps),
               roofline=roofline, kernels=per_kernel)
    if other is not None:
        out["roofline_" + other["bound"]] = other
    out["kernel_timing"] = dict(how=how, kernels_sum_ms=total_s * 1e3, ms_per_step=ms / args.steps,
                                sum_over_step=(total_s * 1e3) / (ms / args.steps) if ms else None,
                                hbm_bound_share=h_share, top_shapes=top_shapes)
    if args.dump_kernels:
        with open(args.dump_kernels, "w") as f:
            f.write("call\tkernel\teager_us\treplay_us\n")
            for sig, kname, ed, d in KERNEL_LIST:
                f.write(f"{sig}\t{kname[:90]}\t{ed:.1f}\t{d:.1f}\n")
    if parity is not None:
        out["parity_checked"] = parity
    if cpu is not None:
        out["cpu_baseline"] = cpu
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    # libraries (NCCL's version banner, torchrun notices) write to fd 1: keep the real stdout for the JSON line only
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="r224", choices=sorted(WORKLOADS))
    ap.add_argument("--math", default="tf32", choices=["fp32", "tf32", "fp32x3"])
    ap.add_argument("--bucket-mb", type=float, default=64.0,
                    help="gradient bucket size of the data-parallel all-reduce. Measured on 8 x B200 (R224, profiles/"
                         "r02_dp8_bucket_*.json): 8 MB buckets overlapped with backward 170.2k img/s, 24 MB 173.7k, one 64 MB "
                         "bucket after backward 174.1k (e2e 146.7k / 154.2k / 171.2k) -- a 45 MB all-reduce takes ~0.13 ms on "
                         "NVSwitch with every SM free and costs more than that in interference when it runs under the "
                         "backward kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-kernel-table", action="store_true")
    ap.add_argument("--dump-kernels", default="", help="write the per-kernel list of one step (launch order, graph-replay us) here")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step (incl. the bucketed all-reduce) as one CUDA graph; auto = on")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()

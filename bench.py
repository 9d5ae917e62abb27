#!/usr/bin/env python
"""bench.py -- ResNet-18 (tests/resnet) training throughput on N B200s, one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W [--workload r56|r224] [--math fp32|tf32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's NumPy algorithm (oracle port) on the host cores

A "step" = zero_grad, forward, cross-entropy, backward, (bucketed all-reduce), fused optimizer step on one
batch of synthetic images through the public xs API of xshinnosuke_b200 (the reference's
tests/resnet/resnet18_dg.py:92-100 loop). `value` keeps the batch resident in HBM; `e2e` uploads every
step's batch from pinned host memory and reads the loss back. See DESIGN.md "Measurement".
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
    "r56": dict(batch=32, hw=56, optimizer="sgd", lr=0.1, classes=100,
                name="ResNet-18 (tests/resnet), batch 32/GPU, synthetic 3x56x56, SGD lr 0.1, dynamic graph"),
    # BASELINE.json configs[4]
    "r224": dict(batch=128, hw=224, optimizer="adam", lr=1e-3, classes=100,
                 name="ResNet-18 (tests/resnet), batch 128/GPU, synthetic 3x224x224, Adam lr 1e-3, dynamic graph"),
}
README_GPU_IMG_S = 265.0   # XS dynamic graph on CuPy, unspecified GPU (README.md:37-38)


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
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


def signature(name, a):
    """Kernel + shape key for the per-launch roofline (conv geometry / GEMM dims / element count)."""
    try:
        if name in ("xsb_conv2d_fwd", "xsb_conv2d_fwd_stats"):
            N, H, W, C, OC, KH, KW, sh, sw, pad = a[4:14]
        elif name in ("xsb_conv2d_dgrad", "xsb_conv2d_wgrad"):
            N, H, W, C, OC, KH, KW, sh, sw, pad = a[3:13]
        elif name == "xsb_gemm":
            return f"{name}[M{a[4]} N{a[5]} K{a[6]}]"
        elif name in ("xsb_bn_stats",):
            return f"{name}[R{a[5]} C{a[6]}]"
        elif name in ("xsb_bn_apply",):
            return f"{name}[R{a[6]} C{a[7]} relu{a[8]}]"
        elif name in ("xsb_bn_bwd", "xsb_bn_train_fwd", "xsb_bn_bwd_dxsum"):
            return f"{name}[R{a[8]} C{a[9]}]"
        elif name == "xsb_colsum":
            return f"{name}[R{a[2]} C{a[3]}]"
        else:
            return name
        return f"{name}[N{N} H{H} W{W} C{C} OC{OC} k{KH}x{KW} s{sh} p{pad}]"
    except Exception:
        return name


# ---------------------------------------------------------------------------------------------- reference arm
def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every host core."""
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=os.cpu_count())
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count()


def run_reference(args, wl):
    """The reference's own algorithm (float64 im2col + np.dot + col2im, oracle/ port of xs/nn/functional.py,
    xs/nn/grad_fn.py, xs/optim) on the host cores, every BLAS thread it can use. Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle.resnet18 import ResNet18 as OracleNet, synthetic_batch
    cores = _use_all_host_threads()
    # bounded sample: a step on the full 224x224 batch is ~1.4 TFLOP of float64 -> time a 4-image slice instead
    n = wl["batch"] if wl["hw"] <= 64 else 4
    x, y = synthetic_batch(n, wl["hw"], wl["classes"], seed=0)
    net = OracleNet(num_classes=wl["classes"])
    state = {}
    step = (lambda: net.train_step(x, y, wl["lr"])) if wl["optimizer"] == "sgd" else \
        (lambda: net.train_step(x, y, wl["lr"], "adam", state))
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = n / dt
    sample = f"{steps} steps of {n} images ({wl['hw']}x{wl['hw']}), float64 NumPy (im2col + np.dot + col2im)"
    out = dict(metric="resnet18_train_images_per_sec", value=val, unit="images/s", n_gpus=args.gpus, steps=steps,
               warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f64", data="synthetic", impl="reference",
               config=dict(workload=wl["name"], sample_batch=n),
               cpu_baseline=dict(value=val, unit="images/s", cores=cores, kind="port", sample=sample),
               e2e=dict(value=val, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(out)


def cpu_baseline_leg(wl, budget_s=20.0):
    from oracle.resnet18 import ResNet18 as OracleNet, synthetic_batch
    cores = _use_all_host_threads()
    n = wl["batch"] if wl["hw"] <= 64 else 2
    x, y = synthetic_batch(n, wl["hw"], wl["classes"], seed=0)
    net = OracleNet(num_classes=wl["classes"])
    state = {}
    step = (lambda: net.train_step(x, y, wl["lr"])) if wl["optimizer"] == "sgd" else \
        (lambda: net.train_step(x, y, wl["lr"], "adam", state))
    step()
    t0, k = time.perf_counter(), 0
    while k < 2 or (time.perf_counter() - t0 < budget_s and k < 12):
        step()
        k += 1
    dt = (time.perf_counter() - t0) / k
    return dict(value=n / dt, unit="images/s", cores=cores, kind="port",
                sample=f"{k} steps of {n} images ({wl['hw']}x{wl['hw']}) through oracle/resnet18.py "
                       f"(float64 NumPy restatement of the reference path), {dt * 1e3:.0f} ms/step")


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args, wl):
    import torch
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.core.device import rt
    from xshinnosuke_b200.recipes import ResNet18
    from xshinnosuke_b200.utils import opcost

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    rt.ensure()
    xs.set_math_mode(args.math)
    dp = None
    B, hw = wl["batch"], wl["hw"]
    rs = np.random.RandomState(1000 + rank)
    x_host = rs.rand(B, 3, hw, hw).astype(np.float32)
    y_host = rs.randint(0, wl["classes"], (B,)).astype(np.int64)
    np.random.seed(0)   # identical initial weights on every rank
    net = ResNet18(num_classes=wl["classes"])
    opt = xs.optim.SGD(net.parameters(), lr=wl["lr"]) if wl["optimizer"] == "sgd" else \
        xs.optim.Adam(net.parameters(), lr=wl["lr"])
    crit = nn.CrossEntropyLoss()
    if world > 1:
        from xshinnosuke_b200.parallel import DataParallel
        dp = DataParallel(net, opt, bucket_mb=args.bucket_mb)
    x_dev, y_dev = xs.tensor(x_host), xs.tensor(y_host)
    x_pin = torch.from_numpy(x_host).pin_memory()
    y_pin = torch.from_numpy(y_host).pin_memory()

    def step_resident():
        opt.zero_grad()
        loss = crit(net(x_dev), y_dev)
        loss.backward()
        opt.step()
        return loss

    def step_e2e():
        # the call a user makes: host batch in, loss value out (resnet18_dg.py:93-100)
        x, y = xs.tensor(x_pin.numpy()), xs.tensor(y_pin.numpy())
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        opt.step()
        return loss.item()

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
    eager_step = step_resident
    graphed = None
    if use_graph:
        # the whole step replayed as one CUDA graph (same kernels, no Python between launches)
        graphed = xs.GraphedStep(eager_step, warmup=max(args.warmup, 3), inputs=(x_dev, y_dev))
        step_resident = graphed
        xh, yh = x_pin.numpy(), y_pin.numpy()
        graphed.prefetch(xh, yh)

        def step_e2e():  # noqa: F811 -- one H2D upload of a pinned host batch and one D2H loss read per step
            loss = graphed(xh, yh)       # staged batch -> captured inputs (device transpose), replay
            graphed.prefetch(xh, yh)     # the NEXT step's upload runs on a side stream under this replay
            return loss.item()

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

    # ---- per-kernel CUDA-event profile (3 extra steps, same stream) for the roofline of the dominant kernel
    prof = {}
    if rank != 0 and dp is not None:
        for _ in range(3):          # the profiled steps contain collectives: every rank must take part
            eager_step()
        torch.cuda.synchronize()
    if rank == 0:
        lib.load()
        names = [n for n in lib.protos if n.startswith("xsb_") and lib.protos[n][1] and
                 n not in ("xsb_init", "xsb_stream_sync", "xsb_memcpy_h2d", "xsb_memcpy_d2h", "xsb_bn_workspace_floats")]
        originals, records = {}, []
        for n in names:
            fn = getattr(lib, n)
            originals[n] = fn

            def wrapped(*a, _fn=fn, _n=n):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                r = _fn(*a)
                e.record()
                records.append((_n, a, s, e))
                return r
            setattr(lib, n, wrapped)
        for _ in range(3):
            eager_step()          # per-call events need the eager path (a graph replay is one opaque launch)
        torch.cuda.synchronize()
        for n, fn in originals.items():
            setattr(lib, n, fn)
        by_shape = {}
        for n, a, s, e in records:
            kind, amount = opcost.cost(n, a)
            ms_ = s.elapsed_time(e)
            d = prof.setdefault(n, dict(ms=0.0, calls=0, amount=0.0, kind=kind))
            d["ms"] += ms_
            d["calls"] += 1
            d["amount"] += amount
            d2 = by_shape.setdefault(signature(n, a), dict(ms=0.0, calls=0, amount=0.0, kind=kind, name=n))
            d2["ms"] += ms_
            d2["calls"] += 1
            d2["amount"] += amount
    if rank != 0:
        return

    pk = peaks()
    total_ms = sum(d["ms"] for d in prof.values()) or 1.0
    # dominant kernel = the (entry point, shape) pair with the most device time: `achieved` is then a true
    # per-launch figure (one shape), not an average over differently sized launches
    tn, td = max(by_shape.items(), key=lambda kv: kv[1]["ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(tn)
    if td["kind"] == "flop":
        achieved = td["amount"] / (td["ms"] * 1e-3) / 1e12
        # kind::tf32 runs at half the bf16 rate; the FFMA (fp32) mode is still reported against the tensor peak
        peak = pk["tc_sustained"]
        roof = dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=traffic,
                    frac_of_tf32_nominal=achieved / (peak / 2.0) if args.math == "tf32" else None,
                    tf32_measured_peak=pk.get("tf32_sustained"),
                    frac_of_tf32_measured=(achieved / pk["tf32_sustained"]) if (args.math == "tf32" and pk.get("tf32_sustained")) else None)
    else:
        achieved = td["amount"] / (td["ms"] * 1e-3) / 1e9
        roof = dict(bound="hbm", achieved=achieved, peak=pk["hbm"], unit="GB/s", frac=achieved / pk["hbm"], traffic=traffic)
    roof.update(kernel=tn, calls_per_step=td["calls"] / 3.0, avg_launch_us=td["ms"] * 1e3 / td["calls"],
                share_of_kernel_time=td["ms"] / total_ms, peak_source=pk["source"],
                algorithmic_per_launch=td["amount"] / td["calls"],
                note="peak = measured dense bf16 (sustained, kernel timed inside a long step); kind::tf32 runs at half "
                     "the bf16 rate, see frac_of_tf32_nominal" if td["kind"] == "flop" else "")
    per_kernel = {n: dict(share=round(d["ms"] / total_ms, 4), us_per_step=round(d["ms"] * 1e3 / 3.0, 1),
                          calls_per_step=round(d["calls"] / 3.0, 1),
                          achieved=round(d["amount"] / (d["ms"] * 1e-3) / (1e12 if d["kind"] == "flop" else 1e9), 2),
                          unit="TFLOP/s" if d["kind"] == "flop" else "GB/s")
                  for n, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    nparams = sum(p.numel() for p in net.parameters())
    value = B * world * args.steps / (ms * 1e-3)
    e2e_val = B * world * args.steps / (ms_e2e * 1e-3)
    cpu = cpu_baseline_leg(wl) if (world == 1 and not args.no_cpu_baseline) else None
    out = dict(metric="resnet18_train_images_per_sec", value=value, unit="images/s", n_gpus=world, steps=args.steps,
               warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
               vs_baseline=value / README_GPU_IMG_S if args.workload == "r56" else None,
               dtype=args.math, data="synthetic",
               config=dict(workload=wl["name"], global_batch=B * world, per_gpu_batch=B, image=f"3x{hw}x{hw}",
                           optimizer=wl["optimizer"], parallelism=f"dp{world}", params=nparams,
                           math_mode=args.math, cuda_graph=bool(use_graph),
                           l2="no flush: per-step working set (activations + 2x45 MB params/grads) exceeds the 126 MB L2",
                           vs_baseline_ref="README.md:37-38 XS dynamic graph on CuPy, 265 img/s, unspecified GPU"),
               clocks=clocks, gpu_launches=launches,
               gpu_launches_unit="C-ABI entry calls of libxsb200.so in the timed region (a lower bound on kernels: a BatchNorm "
                                 "backward is 2 launches, a strided dgrad one per parity class + the filter copy)",
               e2e=dict(value=e2e_val, unit="images/s", h2d_bytes_per_step=int(x_host.nbytes + y_host.nbytes),
                        d2h_bytes_per_step=4, ms_per_step=ms_e2e / args.steps),
               roofline=roof, kernels=per_kernel)
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="r224", choices=sorted(WORKLOADS))
    ap.add_argument("--math", default="tf32", choices=["fp32", "tf32"])
    ap.add_argument("--bucket-mb", type=float, default=8.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

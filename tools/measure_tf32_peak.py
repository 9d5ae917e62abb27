"""Measured dense TF32 tensor-core throughput of this GPU (the denominator for the conv kernels' roofline in tf32 mode).
MEASURED_PEAKS.json holds only the bf16 figure; SURVEY.md 8d asks for the tf32 one measured the same way:
torch.matmul (cuBLAS, allow_tf32) on 8192^3, best of 10 (burst) and back to back for ~3 s (sustained).
Writes profiles/tf32_peak.json. Library GEMM used as a YARDSTICK only -- nothing on the product path calls it."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2 * n ** 3 / (best * 1e-3) / 1e12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(20):
            torch.matmul(a, b)
        k += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = 2 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    out = dict(tf32_tflops=burst, tf32_tflops_sustained=sustained, how="torch.matmul allow_tf32 8192^3 (cuBLAS): best of 10 "
               "(burst), back to back for 3 s (sustained)", gpu=torch.cuda.get_device_name(0),
               theoretical_tflops_at_max_clock=148 * 2048 * 2 * 1.965e9 / 1e12,
               theoretical_note="tools/mma_rate.cu: tcgen05.mma kind::tf32 issues 2048 MAC/clk/SM (N >= 128); x 148 SMs x 1.965 GHz")
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "tf32_peak.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

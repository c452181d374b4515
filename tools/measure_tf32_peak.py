"""TF32 tensor-core peak of this GPU, measured the way the driver measures BF16 in MEASURED_PEAKS.json:
torch.matmul (cuBLAS, fp32 operands with allow_tf32) 8192^3, 2*N^3 flops, best of 10 (burst) and back to back for
4 s (sustained).  Writes profiles/tf32_peak.json (read by bench.py for the alignment contraction's roofline).
Usage (GPU box): python tools/measure_tf32_peak.py"""
import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.backends.cuda.matmul.allow_tf32 = True
N = 8192
a = torch.randn(N, N, device="cuda")
b = torch.randn(N, N, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = None
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a @ b
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    best = t if best is None else min(best, t)
burst = 2.0 * N ** 3 / (best * 1e-3) / 1e12
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0, n = time.time(), 0
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(20):
        a @ b
    n += 20
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
sustained = 2.0 * N ** 3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e12
out = {"tf32_tflops": burst, "tf32_tflops_sustained": sustained, "gpu_name": torch.cuda.get_device_name(0),
       "torch": torch.__version__, "how": "torch.matmul fp32 with allow_tf32 (cuBLAS TF32) 8192^3 (2*N^3): best of 10 "
       "(burst) and back to back for 4 s (sustained), CUDA events", "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tf32_peak.json"), "w"), indent=1)
print(json.dumps(out))

"""Measures the dense TF32 tensor-pipe peak the way MEASURED_PEAKS.json's bf16 figure was taken (torch.matmul 8192^3,
best of 10 and a 4 s sustained loop): the roofline denominator for the 3xTF32 contraction is this / 3."""
import json
import time

import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
k = 0
while time.time() - t0 < 4.0:
    for _ in range(10):
        a @ b
    k += 10
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
sustained = e0.elapsed_time(e1) / k
out = {"tf32_tflops": 2 * n**3 / best / 1e9, "tf32_tflops_sustained": 2 * n**3 / sustained / 1e9,
       "how": "torch.matmul fp32 with allow_tf32, 8192^3, best of 10 (burst) and back to back for 4 s (sustained)"}
out["x3_tf32_peak_tflops"] = out["tf32_tflops"] / 3
out["x3_tf32_peak_tflops_sustained"] = out["tf32_tflops_sustained"] / 3
print(json.dumps(out))

#!/usr/bin/env python
"""cuBLAS bf16 8192^3 matmul — the launch MEASURED_PEAKS.json's tensor peak is defined by — so that ncu's
`sm__pipe_tensor_cycles_active` can be read for it under exactly the settings used for this repo's kernels
(tools/gpu_suite.sh ncucalib).  Prints the CUDA-event TFLOP/s of the same launches (never taken under ncu)."""
import json
import torch

n = 8192
a = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
b = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    c = a @ b
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"what": "torch.matmul bf16 8192^3 (cuBLAS)", "ms": ms, "tflops": 2 * n ** 3 / ms / 1e9}))

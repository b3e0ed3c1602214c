#!/usr/bin/env python
"""A few MLP forwards of one 75 776-row chunk: the launch list / ncu target for the two GEMM kernels."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from kubetorch_b200.device import mlp, ops  # noqa: E402

ops.ensure_init([0])
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = torch.Generator(device="cuda").manual_seed(0)
w1 = (torch.randn(1024, 256, device="cuda", generator=g) * 0.02).bfloat16()
w2 = (torch.randn(1024, 1024, device="cuda", generator=g) * 0.02).bfloat16()
w3 = (torch.randn(64, 1024, device="cuda", generator=g) * 0.02).bfloat16()
obs = torch.randn(rows, 256, device="cuda", generator=g).bfloat16()
y = mlp.mlp_forward(obs, w1, w2, w3)
for _ in range(reps):
    mlp.mlp_forward(obs, w1, w2, w3, out=y)
torch.cuda.synchronize()
print("ok", float(y.float().abs().sum()))

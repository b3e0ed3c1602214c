#!/usr/bin/env python
"""Runs each hot kernel a few times at its benchmark size, for `ncu -k regex:...` captures."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from kubetorch_b200.device import mlp, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ops.ensure_init([0])
n = 1 << 26
x = torch.randn(n, device="cuda")
y = torch.empty_like(x)
if which in ("all", "map"):
    for _ in range(4):
        ops.map_tensor(x, "scale", 2.0, out=y)
if which in ("all", "tma"):
    for _ in range(4):
        ops.map_tensor(x, "scale", 2.0, out=y, variant=2)
if which in ("all", "reduce"):
    for _ in range(4):
        ops.map_reduce_sum(x, "scale", 2.0)
if which in ("all", "pack"):
    ts = [torch.empty(1 << 18, dtype=torch.uint8, device="cuda") for _ in range(4096)]   # C4's shard list, 1 GiB
    plan = ops.PackPlan(ts)
    for _ in range(4):
        plan.run()
if which in ("all", "gemm"):
    g = torch.Generator(device="cuda").manual_seed(0)
    w1 = (torch.randn(1024, 256, device="cuda", generator=g) * 0.02).bfloat16()
    w2 = (torch.randn(1024, 1024, device="cuda", generator=g) * 0.02).bfloat16()
    w3 = (torch.randn(64, 1024, device="cuda", generator=g) * 0.02).bfloat16()
    obs = torch.randn(75776, 256, device="cuda", generator=g).bfloat16()   # one default chunk: 4 waves of 74 CTA pairs
    for _ in range(3):
        mlp.mlp_forward(obs, w1, w2, w3)
torch.cuda.synchronize()
print("done", which)

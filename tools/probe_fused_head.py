#!/usr/bin/env python
"""Bring-up probe for the fused layer-2 + head kernel (ktb_set_tuning key 18): compare with the unfused pair kernel
(fp32 partial sums are added in a different order, so compare within bf16 rounding), then time both.
Run under `timeout`: a protocol error traps (bounded mbarrier waits) rather than hangs."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from kubetorch_b200.device import mlp, ops  # noqa: E402
from tools.bench_mlp import timeit  # noqa: E402

ops.ensure_init([0])
g = torch.Generator(device="cuda").manual_seed(0)
w1 = (torch.randn(1024, 256, device="cuda", generator=g) * 0.02).bfloat16()
w2 = (torch.randn(1024, 1024, device="cuda", generator=g) * 0.02).bfloat16()
w3 = (torch.randn(64, 1024, device="cuda", generator=g) * 0.02).bfloat16()
flop_per_row = 2 * (256 * 1024 + 1024 * 1024 + 1024 * 64)
for M in (256, 4096, 75776, 262144):
    obs = torch.randn(M, 256, device="cuda", generator=g).bfloat16()
    ops.set_tuning(18, 0)
    ops.set_tuning(8, 65536)
    ref = mlp.mlp_forward(obs, w1, w2, w3).clone()
    torch.cuda.synchronize()
    ops.set_tuning(18, 1)
    out = mlp.mlp_forward(obs, w1, w2, w3)
    torch.cuda.synchronize()
    diff = (out.float() - ref.float()).abs()
    tol = 2.0 ** -7 * ref.float().abs() + 1e-3
    row = {"M": M, "max_abs_diff": diff.max().item(), "mismatching_elems": int((out != ref).sum().item()),
           "outside_bf16_rounding": int((diff > tol).sum().item()), "numel": out.numel()}
    ok = row["outside_bf16_rounding"] == 0
    if ok and M >= 75776:
        for fused, chunk in ((0, 65536), (1, 65536), (1, 75776), (1, 37888)):
            ops.set_tuning(18, fused)
            ops.set_tuning(8, chunk)
            mlp._scratch.clear()
            ms = timeit(lambda: mlp.mlp_forward(obs, w1, w2, w3, out=out))
            row[f"fused{fused}_chunk{chunk}_tflops"] = round(flop_per_row * M / ms / 1e9, 1)
    print(json.dumps(row), flush=True)
    if not ok:
        break
ops.set_tuning(18, 0)
ops.set_tuning(8, 65536)

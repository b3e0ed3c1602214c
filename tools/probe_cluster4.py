#!/usr/bin/env python
"""Bring-up probe for the cluster-of-4 multicast GEMM (ktb_set_tuning key 15): bit-compare against the CTA-pair
kernel, then time both.  Run under `timeout`: a protocol error traps (bounded mbarrier waits) rather than hangs."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from kubetorch_b200.device import lib as L  # noqa: E402
from kubetorch_b200.device import mlp, ops  # noqa: E402
from tools.bench_mlp import timeit  # noqa: E402

ops.ensure_init([0])
g = torch.Generator(device="cuda").manual_seed(0)
w1 = (torch.randn(1024, 256, device="cuda", generator=g) * 0.02).bfloat16()
w2 = (torch.randn(1024, 1024, device="cuda", generator=g) * 0.02).bfloat16()
w3 = (torch.randn(64, 1024, device="cuda", generator=g) * 0.02).bfloat16()
flop_per_row = 2 * (256 * 1024 + 1024 * 1024 + 1024 * 64)
for M in (4096, 262144):
    obs = torch.randn(M, 256, device="cuda", generator=g).bfloat16()
    ops.set_tuning(15, 0)
    ref = mlp.mlp_forward(obs, w1, w2, w3).clone()
    torch.cuda.synchronize()
    ops.set_tuning(15, 1)
    out = mlp.mlp_forward(obs, w1, w2, w3)
    torch.cuda.synchronize()
    same = bool(torch.equal(out, ref))
    row = {"M": M, "bit_identical_to_pair_kernel": same,
           "max_abs_diff": (out.float() - ref.float()).abs().max().item()}
    if same and M >= 65536:
        for key in (0, 1):
            ops.set_tuning(15, key)
            ms = timeit(lambda: mlp.mlp_forward(obs, w1, w2, w3, out=out))
            row["cluster4_ms" if key else "pair_ms"] = ms
            row["cluster4_tflops" if key else "pair_tflops"] = flop_per_row * M / ms / 1e9
        ops.set_tuning(15, 0)
        ops.set_tuning(17, 5)
        out5 = mlp.mlp_forward(obs, w1, w2, w3)
        row["stages5_bit_identical"] = bool(torch.equal(out5, ref))
        ms = timeit(lambda: mlp.mlp_forward(obs, w1, w2, w3, out=out))
        row["pair_stages5_ms"], row["pair_stages5_tflops"] = ms, flop_per_row * M / ms / 1e9
        ops.set_tuning(17, 4)
    row["max_active_clusters_of_4"] = L.call("ktb_set_tuning", 16, 0)
    print(json.dumps(row), flush=True)
    if not same:
        break
ops.set_tuning(15, 0)

#!/usr/bin/env python
"""Layer-1 B-resident kernel (ktb_set_tuning(24, 1)) against the shipped pair kernel: bit-identity and time."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from kubetorch_b200.device import mlp, ops  # noqa: E402

ops.ensure_init([0])
g = torch.Generator(device="cuda").manual_seed(0)
w1 = (torch.randn(1024, 256, device="cuda", generator=g) * 0.02).bfloat16()
w2 = (torch.randn(1024, 1024, device="cuda", generator=g) * 0.02).bfloat16()
w3 = (torch.randn(64, 1024, device="cuda", generator=g) * 0.02).bfloat16()


def run(rows, iters=10):
    obs = torch.randn(rows, 256, device="cuda", generator=g).bfloat16()
    out = {}
    for flag in (0, 1, 2, 3, 4):
        ops.set_tuning(24, flag)
        y = mlp.mlp_forward(obs, w1, w2, w3)
        torch.cuda.synchronize()
        for _ in range(3):
            mlp.mlp_forward(obs, w1, w2, w3, out=y)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            mlp.mlp_forward(obs, w1, w2, w3, out=y)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        out[flag] = (y.clone(), ms)
    ops.set_tuning(24, 2)
    same = bool(torch.equal(out[0][0], out[1][0])) and bool(torch.equal(out[0][0], out[2][0])) and bool(torch.equal(out[0][0], out[3][0])) and bool(torch.equal(out[0][0], out[4][0]))
    flop = 2 * (256 * 1024 + 1024 * 1024 + 1024 * 64) * rows
    print(json.dumps({"rows": rows, "bit_identical": same, "ms_pair_kernel": out[0][1], "ms_bres_kernel": out[1][1], "ms_warp_store": out[2][1], "ms_warp_store_8w": out[3][1], "ms_quarter_boxes_8st": out[4][1], "tflops_quarter_boxes_8st": flop / out[4][1] / 1e9, "tflops_8w": flop / out[3][1] / 1e9, "tflops_warp_store": flop / out[2][1] / 1e9,
                      "tflops_pair": flop / out[0][1] / 1e9, "tflops_bres": flop / out[1][1] / 1e9}), flush=True)
    return same


ok = True
for rows in (256, 512, 1280, 75776, 262144, 75776 + 256 * 7):
    ok = run(rows) and ok
print("ALL BIT-IDENTICAL" if ok else "MISMATCH")

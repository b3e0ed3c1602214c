#!/usr/bin/env python
"""Generic MLP tuning probe: ktb_set_tuning(KEY, mode) for KEY, MODES from argv (default key 25).
Epilogue hand-back semantics (ktb_set_tuning(25, mode)) for the layer-1 and the fused layer-2+head kernels:
bit-identity against mode 0 and time.  mode & 3: 0 = release.cluster arrives, 1 = CTA-scope release for the TMEM
hand-backs, 2 = also for c_ready; mode & 4: software-pipelined TMEM loads in the fused kernel's epilogue."""
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
KEY = int(sys.argv[1]) if len(sys.argv) > 1 else 25
MODES = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (0, 1, 2, 5, 6)
RESET = int(sys.argv[3]) if len(sys.argv) > 3 else MODES[0]


def run(rows, iters=10, reps=3):
    obs = torch.randn(rows, 256, device="cuda", generator=g).bfloat16()
    out = {}
    for mode in MODES:
        ops.set_tuning(KEY, mode)
        y = mlp.mlp_forward(obs, w1, w2, w3)
        torch.cuda.synchronize()
        for _ in range(3):
            mlp.mlp_forward(obs, w1, w2, w3, out=y)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                mlp.mlp_forward(obs, w1, w2, w3, out=y)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / iters)
        out[mode] = (y.clone(), best)
    ops.set_tuning(KEY, RESET)
    same = all(bool(torch.equal(out[MODES[0]][0], out[m][0])) for m in MODES)
    flop = 2 * (256 * 1024 + 1024 * 1024 + 1024 * 64) * rows
    print(json.dumps({"rows": rows, "bit_identical": same,
                      "ms": {str(m): round(out[m][1], 5) for m in MODES},
                      "tflops": {str(m): round(flop / out[m][1] / 1e9, 1) for m in MODES}}), flush=True)
    return same


ok = True
for rows in (256, 512, 1280, 75776, 262144, 75776 + 256 * 7, 262144):
    ok = run(rows) and ok
print("ALL BIT-IDENTICAL" if ok else "MISMATCH")

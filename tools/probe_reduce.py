#!/usr/bin/env python
"""Where the fixed cost of map_reduce_sum goes: size sweep with and without the cross-CTA fold."""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from kubetorch_b200.device import ops
from tools.sweep_pack_reduce import timeit

ops.ensure_init([0])
for mib in (128, 256, 512, 1024, 2048):
    x = torch.randn(mib << 18, device="cuda")
    row = {"mib": mib}
    for fold in (1, 0):
        ops.set_tuning(14, fold)
        ms = timeit(lambda: ops.map_reduce_sum(x, "scale", 2.0), iters=30)
        row["fold%d_us" % fold] = round(ms * 1e3, 2)
    ops.set_tuning(14, 1)
    y = torch.empty_like(x)
    ms = timeit(lambda: ops.map_tensor(x, "scale", 2.0, out=y), iters=30)
    row["map_us"] = round(ms * 1e3, 2)
    print(json.dumps(row), flush=True)
    del x, y

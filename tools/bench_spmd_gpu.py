#!/usr/bin/env python
"""Arbitrary Python callable (NOT @mapped) on N GPU rank processes: CUDA tensor args/results travel through HBM
arenas (ktb_pack → ktb_broadcast → zero-copy views → ktb_pack → ktb_unpack). Host-clock per-call time."""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.environ["PYTHONPATH"] = os.pathsep.join([REPO] + [p for p in os.environ.get("PYTHONPATH", "").split(os.pathsep) if p])
import torch  # noqa: E402

import kubetorch_b200 as kt  # noqa: E402
from oracle import cases  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    remote = kt.fn(cases.double, name="spmd-gpu-bench").to(
        kt.Compute(gpus=n, allowed_serialization=["json", "pickle"]).distribute("spmd", workers=1, num_proc=n,
                                                                                 arena_bytes=300 << 20))
    for log2 in (10, 20, 24, 26, 28):
        nb = 1 << log2
        x = torch.randn(nb // 4, device="cuda:0")
        for _ in range(3):
            out = remote(x, serialization="pickle")
        iters = 200 if log2 <= 20 else 20
        t0 = time.perf_counter()
        for _ in range(iters):
            out = remote(x, serialization="pickle")
        torch.cuda.synchronize(0)
        dt = (time.perf_counter() - t0) / iters
        ok = bool(torch.equal(torch.cat(out), x * 2))
        print(json.dumps({"what": "spmd_gpu_arbitrary_callable", "n_ranks": n, "log2_bytes": log2, "ms": dt * 1e3,
                          "calls_per_sec": 1 / dt, "arg_plus_result_gbps": 2 * nb / dt / 1e9, "ok": ok}), flush=True)
    remote.teardown()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Small calls through rank PROCESSES on the CPU backends (no GPU): our pool vs the reference port (oracle), same callable,
same payload.  Best of several batches (shared boxes are noisy)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import kubetorch_b200 as kt  # noqa: E402
from oracle import cases  # noqa: E402
from oracle.ref_dispatch import OracleRuntime  # noqa: E402


def best_rate(fn, n=200, batches=7, warm=20):
    for _ in range(warm):
        fn()
    best = 0.0
    for _ in range(batches):
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        best = max(best, n / (time.perf_counter() - t0))
    return best


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    for nbytes in (1 << 10, 1 << 20):
        x = torch.randn(nbytes // 4)
        f = kt.fn(cases.identity, name=f"small-{nbytes}").to(
            kt.Compute(cpus="1", allowed_serialization=["json", "pickle"]).distribute("spmd", workers=1, num_proc=P))
        ours = best_rate(lambda: f(x, serialization="pickle"))
        f.teardown()
        rt = OracleRuntime("oracle.cases", "identity", P)
        ref = best_rate(lambda: rt.call(x, serialization="pickle"))
        rt.close()
        print(f"identity {nbytes} B, {P} ranks: ours {ours:.0f} calls/s, reference port {ref:.0f} calls/s "
              f"({ours / ref:.1f}x)", flush=True)


if __name__ == "__main__":
    main()

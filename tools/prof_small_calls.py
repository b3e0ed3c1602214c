#!/usr/bin/env python
"""Host-side cost of one small remote call through the public API: calls/s and a cProfile of the call path.
`--gpu` profiles remote(x_1KiB) on the b200 backend (1 GPU); default profiles C1 hello_world in-process (CPU)."""
import argparse
import cProfile
import os
import pstats
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import kubetorch_b200 as kt  # noqa: E402
from oracle import cases  # noqa: E402  (bench/test tooling only)


def rate(fn, n):
    for _ in range(200):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return n / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("-n", type=int, default=20000)
    a = ap.parse_args()
    if a.gpu:
        import types

        import torch
        src = cases.double
        double = types.FunctionType(src.__code__, src.__globals__, "double_prof", src.__defaults__, src.__closure__)
        double = kt.mapped("scale", alpha=2.0)(double)
        remote = kt.fn(double, name="prof-double").to(
            kt.Compute(gpus=a.gpus).distribute("b200", workers=1, num_proc=a.gpus))
        x = torch.randn(256, device="cuda:0")
        call = lambda: remote(x, serialization="pickle")  # noqa: E731
        sync = torch.cuda.synchronize
    else:
        remote = kt.fn(cases.hello_world).to(kt.Compute(cpus=".1"))
        call = lambda: remote()  # noqa: E731
        sync = lambda: None  # noqa: E731
    print("calls/s:", round(rate(call, a.n)))
    sync()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5000):
        call()
    pr.disable()
    sync()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
    remote.teardown()


if __name__ == "__main__":
    main()

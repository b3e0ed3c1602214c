#!/usr/bin/env python
"""BASELINE config C5: identity map over tensors of 1 KB – 1 GB at 1/2/4/8 GPUs (device-resident on GPU 0,
device-timed) next to the reference CPU dispatch (oracle port) on the same box. JSON lines →
gpurun_out/sweep_c5.jsonl.  Usage: python tools/sweep_c5.py [max_gpus] [--no-ref]"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import kubetorch_b200 as kt  # noqa: E402
from kubetorch_b200.device import ops  # noqa: E402
from oracle import cases  # noqa: E402

OUT = os.path.join(REPO, "gpurun_out", "sweep_c5.jsonl")
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    with open(OUT, "a") as f:
        f.write(line + "\n")


def _clone(fn):
    import types

    return types.FunctionType(fn.__code__, fn.__globals__, fn.__name__, fn.__defaults__, fn.__closure__)


def main():
    max_gpus = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else torch.cuda.device_count()
    do_ref = "--no-ref" not in sys.argv
    for n_gpus in [n for n in (1, 2, 4, 8) if n <= max_gpus]:
        devices = list(range(n_gpus))
        ops.ensure_init(devices)
        for k in range(10, 31, 2):
            nb = 1 << k
            x = torch.empty(nb, dtype=torch.uint8, device="cuda:0")
            y = torch.empty_like(x)
            iters = 200 if k <= 20 else (50 if k <= 26 else 10)
            modes = {"fused_pull_push": lambda: ops.scatter_map_gather(x, "identity", devices=devices, out_root=y)}
            if n_gpus > 1 and nb >= (1 << 23):
                sess = ops.PushSession(devices, ops.shard_bounds(nb, n_gpus, 0)[1])
                modes["push_pipeline"] = lambda: sess.call(x, y, "identity")
            best = None
            for name, fn in modes.items():
                for _ in range(3):
                    fn()
                for d in devices:
                    torch.cuda.synchronize(d)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    fn()
                e1.record()
                for d in devices:
                    torch.cuda.synchronize(d)
                ms = e0.elapsed_time(e1) / iters
                if best is None or ms < best[1]:
                    best = (name, ms)
            emit(what="c5_identity", n_gpus=n_gpus, log2_bytes=k, transfer=best[0], ms=best[1],
                 calls_per_sec=1e3 / best[1], arg_plus_result_gbps=2 * nb / best[1] / 1e6)
            del x, y
    # public-API call rate for small payloads (Python __call__ overhead included, host clock)
    ident = kt.mapped("identity")(_clone(cases.identity))
    remote = kt.fn(ident, name="c5-api").to(kt.Compute(gpus=1).distribute("b200", workers=1, num_proc=1))
    xs = torch.empty(1024, dtype=torch.uint8, device="cuda:0")
    for _ in range(100):
        remote(xs, serialization="pickle")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5000):
        remote(xs, serialization="pickle")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    emit(what="public_api_1KiB_calls", calls_per_sec=5000 / dt, us_per_call=dt / 5000 * 1e6)
    remote.teardown()
    hw = kt.fn(cases.hello_world, name="c1-hello").to(kt.Compute(cpus=".1"))
    t0 = time.perf_counter()
    for _ in range(20000):
        hw()
    dt = time.perf_counter() - t0
    emit(what="c1_hello_world_in_process", calls_per_sec=20000 / dt)
    if do_ref:
        from oracle.ref_dispatch import OracleRuntime

        with OracleRuntime("oracle.cases", "identity", 8, "spmd", extra_path=REPO) as rt:
            for k in range(10, 25, 2):
                x = torch.empty(1 << k, dtype=torch.uint8)
                rt.call(x, serialization="pickle")
                n = 50 if k <= 16 else (10 if k <= 20 else 3)
                t0 = time.perf_counter()
                for _ in range(n):
                    rt.call(x, serialization="pickle")
                dt = (time.perf_counter() - t0) / n
                emit(what="c5_reference_cpu", ranks=8, log2_bytes=k, ms=dt * 1e3, calls_per_sec=1 / dt,
                     arg_plus_result_gbps=2 * (1 << k) / dt / 1e9)
        with OracleRuntime("oracle.cases", "hello_world", 1, "spmd", extra_path=REPO) as rt:
            rt.call(serialization="json")
            t0 = time.perf_counter()
            for _ in range(300):
                rt.call(serialization="json")
            dt = time.perf_counter() - t0
            emit(what="c1_hello_world_reference_port", calls_per_sec=300 / dt)


if __name__ == "__main__":
    main()

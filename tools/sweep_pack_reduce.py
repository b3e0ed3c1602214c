#!/usr/bin/env python
"""Pack / batched-call / reduce sweep on one B200 (descriptor batch size, reduce tile, torch references).
Writes JSON lines to gpurun_out/sweep_pack_reduce.jsonl."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from kubetorch_b200.device import ops  # noqa: E402

OUT = os.path.join(REPO, "gpurun_out", "sweep_pack_reduce.jsonl")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
out_f = open(OUT, "a")


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    out_f.write(line + "\n")
    out_f.flush()


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters  # ms


def main():
    ops.ensure_init([0])
    # reduce: tile (loads in flight), size, torch.sum beside it
    for n in (1 << 26, 1 << 28):
        x = torch.randn(n, device="cuda")
        nbytes = n * 4
        for loads in (4, 8):
            ops.set_tuning(13, loads)
            ms = timeit(lambda: ops.map_reduce_sum(x, "scale", 2.0))
            emit(what="reduce_sum_f32", mib=nbytes >> 20, loads=loads, ms=ms, gbps=nbytes / ms / 1e6)
        ops.set_tuning(13, 4)
        ms = timeit(lambda: torch.sum(x))
        emit(what="torch_sum_f32", mib=nbytes >> 20, ms=ms, gbps=nbytes / ms / 1e6)
        xb = x.view(torch.bfloat16)
        for loads in (4, 8):
            ops.set_tuning(13, loads)
            ms = timeit(lambda: ops.map_reduce_sum(xb, "identity"))
            emit(what="reduce_sum_bf16", mib=nbytes >> 20, loads=loads, ms=ms, gbps=nbytes / ms / 1e6)
        ops.set_tuning(13, 4)
        del x, xb

    # pack: C4's shard list (4096 x 256 KiB is 1 GiB; 1024 x 256 KiB kept for continuity), small leaves, big leaves
    for cnt, sz in ((1024, 1 << 18), (4096, 1 << 18), (4096, 1 << 12), (8, 1 << 25)):
        ts = [torch.empty(sz, dtype=torch.uint8, device="cuda") for _ in range(cnt)]
        plan = ops.PackPlan(ts)
        for large in (0, 1):
            ops.set_tuning(12, large)
            ms = timeit(lambda: plan.run(), iters=10)
            emit(what="pack_plan", count=cnt, seg_bytes=sz, large_params=large, ms=ms, gbps=2 * cnt * sz / ms / 1e6)
        cat_ms = timeit(lambda: torch.cat(ts), iters=10)
        emit(what="torch_cat", count=cnt, seg_bytes=sz, ms=cat_ms, gbps=2 * cnt * sz / cat_ms / 1e6)
        del ts, plan

    # batched small calls: 4096 x 1 KiB mapped calls in one plan
    xs = [torch.randn(256, device="cuda") for _ in range(4096)]
    outs = [torch.empty_like(t) for t in xs]
    bplan = ops.BatchPlan(xs, outs, "scale", 2.0)
    for large in (0, 1):
        ops.set_tuning(12, large)
        ms = timeit(lambda: bplan.run(), iters=20)
        emit(what="map_batch_plan_4096x1KiB", large_params=large, ms=ms, calls_per_sec=4096 / ms * 1e3)
    ok = all(torch.equal(o, x * 2) for x, o in zip(xs[:64], outs[:64]))
    emit(what="map_batch_plan_check", ok=bool(ok))


if __name__ == "__main__":
    main()

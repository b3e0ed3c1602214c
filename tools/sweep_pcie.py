#!/usr/bin/env python
"""PCIe microbenchmarks on one GPU: what each direction gives alone / together, and hybrid forms of the
host-resident call (copy engine one way, kernel zero-copy the other)."""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from kubetorch_b200.device import lib as L  # noqa: E402
from kubetorch_b200.device import ops  # noqa: E402


def wall(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters


def main():
    ops.ensure_init([0])
    n = 1 << 26
    nb = n * 4
    xh = torch.randn(n).pin_memory()
    yh = torch.empty(n).pin_memory()
    xd = torch.empty(n, device="cuda")
    yd = torch.empty(n, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = []

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    dt = wall(lambda: xd.copy_(xh, non_blocking=True))
    emit(what="h2d_alone", gbps=nb / dt / 1e9)
    dt = wall(lambda: yh.copy_(yd, non_blocking=True))
    emit(what="d2h_alone", gbps=nb / dt / 1e9)

    def duplex():
        with torch.cuda.stream(s1):
            xd.copy_(xh, non_blocking=True)
        with torch.cuda.stream(s2):
            yh.copy_(yd, non_blocking=True)
    dt = wall(duplex)
    emit(what="h2d_and_d2h_together", gbps_each_way=nb / dt / 1e9)
    st = torch.cuda.current_stream().cuda_stream

    def k_read_host():   # kernel loads host memory over PCIe, stores to HBM
        L.call("ktb_map", 0, L.OP_SCALE, L.F32, xh.data_ptr(), yd.data_ptr(), n, 2.0, 0.0, L.VARIANT_VEC, st)
    dt = wall(k_read_host)
    emit(what="kernel_reads_host", gbps=nb / dt / 1e9)

    def k_write_host():  # kernel loads HBM, stores to host memory over PCIe
        L.call("ktb_map", 0, L.OP_SCALE, L.F32, xd.data_ptr(), yh.data_ptr(), n, 2.0, 0.0, L.VARIANT_VEC, st)
    dt = wall(k_write_host)
    emit(what="kernel_writes_host", gbps=nb / dt / 1e9)

    # hybrid A: copy engine H2D in chunks, kernel writes results straight to host (no D2H copies, no staging out)
    for chunk in (4 << 20, 16 << 20):
        ce = chunk // 4
        stage = [torch.empty(ce, device="cuda") for _ in range(2)]
        evs = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]

        def hybrid_a():
            for i, off in enumerate(range(0, n, ce)):
                b = i & 1
                m = min(ce, n - off)
                with torch.cuda.stream(s1):
                    if i >= 2:
                        s1.wait_event(done[b])
                    stage[b][:m].copy_(xh[off:off + m], non_blocking=True)
                    evs[b].record(s1)
                with torch.cuda.stream(s2):
                    s2.wait_event(evs[b])
                    L.call("ktb_map", 0, L.OP_SCALE, L.F32, stage[b].data_ptr(), yh.data_ptr() + off * 4, m, 2.0, 0.0,
                           L.VARIANT_VEC, s2.cuda_stream)
                    done[b].record(s2)
        dt = wall(hybrid_a, iters=3)
        ok = bool(torch.equal(yh[-1000:], xh[-1000:] * 2))
        emit(what="hybrid_h2d_copy_kernel_writes_host", chunk=chunk, arg_plus_result_gbps=2 * nb / dt / 1e9, ok=ok)

    # hybrid B: kernel reads host (zero-copy), results D2H by copy engine in chunks
    for chunk in (4 << 20, 16 << 20):
        ce = chunk // 4
        stage = [torch.empty(ce, device="cuda") for _ in range(2)]
        evs = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]

        def hybrid_b():
            for i, off in enumerate(range(0, n, ce)):
                b = i & 1
                m = min(ce, n - off)
                with torch.cuda.stream(s1):
                    if i >= 2:
                        s1.wait_event(done[b])
                    L.call("ktb_map", 0, L.OP_SCALE, L.F32, xh.data_ptr() + off * 4, stage[b].data_ptr(), m, 2.0, 0.0,
                           L.VARIANT_VEC, s1.cuda_stream)
                    evs[b].record(s1)
                with torch.cuda.stream(s2):
                    s2.wait_event(evs[b])
                    yh[off:off + m].copy_(stage[b][:m], non_blocking=True)
                    done[b].record(s2)
        dt = wall(hybrid_b, iters=3)
        ok = bool(torch.equal(yh[-1000:], xh[-1000:] * 2))
        emit(what="hybrid_kernel_reads_host_d2h_copy", chunk=chunk, arg_plus_result_gbps=2 * nb / dt / 1e9, ok=ok)

    dt = wall(lambda: ops.map_host(xh, "scale", 2.0, out_host=yh, chunk_bytes=16 << 20), iters=3)
    emit(what="ktb_map_host_pipelined_16MiB", arg_plus_result_gbps=2 * nb / dt / 1e9)


if __name__ == "__main__":
    main()

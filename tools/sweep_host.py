#!/usr/bin/env python
"""Host-resident e2e experiments on N GPUs: public API with (a) an ordinary pinned tensor and (b) a pinned
tensor whose shard pages were first-touched on each GPU's NUMA node. JSON lines → gpurun_out/sweep_host.jsonl"""
import ctypes
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402


def _clone(fn):
    """A copy of an oracle callable to decorate (the shared function object stays undecorated)."""
    import types

    return types.FunctionType(fn.__code__, fn.__globals__, fn.__name__, fn.__defaults__, fn.__closure__)

import kubetorch_b200 as kt  # noqa: E402
from kubetorch_b200.device import ops  # noqa: E402
from oracle import cases  # noqa: E402

OUT = os.path.join(REPO, "gpurun_out", "sweep_host.jsonl")
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    with open(OUT, "a") as f:
        f.write(line + "\n")


def gpu_numa_cpus(dev):
    """CPUs local to GPU `dev` (sysfs), or None."""
    try:
        bdf = torch.cuda.get_device_properties(dev).pci_bus_id if hasattr(torch.cuda.get_device_properties(dev), "pci_bus_id") else None
    except Exception:  # noqa: BLE001
        bdf = None
    try:
        import subprocess

        q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(dev)],
                           capture_output=True, text=True).stdout.strip().lower()
        bdf = q[4:] if q.startswith("0000") and len(q) > 12 else q
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        out = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            out.update(range(int(a), int(b or a) + 1))
        return out
    except Exception:  # noqa: BLE001
        return None


def numa_pinned(n_elems, n_gpus):
    """One contiguous fp32 tensor; shard r's pages first-touched from a thread bound to GPU r's NUMA node; then
    page-locked in place with cudaHostRegister."""
    x = torch.empty(n_elems, dtype=torch.float32)
    allowed = os.sched_getaffinity(0)

    def touch(r):
        cpus = gpu_numa_cpus(r)
        if cpus:
            try:
                os.sched_setaffinity(0, cpus & allowed or allowed)
            except OSError:
                pass
        b, e = ops.shard_bounds(n_elems, n_gpus, r)
        x[b:e].normal_()

    ths = [threading.Thread(target=touch, args=(r,)) for r in range(n_gpus)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    rc = torch.cuda.cudart().cudaHostRegister(x.data_ptr(), x.numel() * 4, 0)
    assert int(rc) == 0, rc
    return x


def main():
    n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    n = 1 << 26
    double = kt.mapped("scale", alpha=2.0)(_clone(cases.double))
    emit(what="numa", cpus={r: (sorted(gpu_numa_cpus(r))[:2] if gpu_numa_cpus(r) else None) for r in range(n_gpus)})
    bufs = {"plain_pinned": torch.randn(n).pin_memory(), "numa_first_touch_registered": numa_pinned(n, n_gpus)}
    for host_mode in ("threads", "multi"):
        remote = kt.fn(double, name=f"host-sweep-{host_mode}").to(
            kt.Compute(gpus=n_gpus).distribute("b200", workers=1, num_proc=n_gpus, host_mode=host_mode))
        for kind, xh in bufs.items():
            assert xh.is_pinned()
            for _ in range(2):
                out = remote(xh, serialization="pickle")
            t0 = time.perf_counter()
            for _ in range(8):
                out = remote(xh, serialization="pickle")
            dt = (time.perf_counter() - t0) / 8
            ok = bool(torch.equal(torch.cat(out)[-4096:], xh[-4096:] * 2))
            emit(what="e2e_host", host_mode=host_mode, kind=kind, n_gpus=n_gpus, ms=dt * 1e3,
                 gbps=2 * n * 4 / dt / 1e9, ok=ok)
        remote.teardown()
    # raw pipeline without the API layer, by chunk size
    from kubetorch_b200.device import ops as _ops

    xh = bufs["plain_pinned"]
    yh = torch.empty_like(xh).pin_memory()
    for cb in (2 << 20, 4 << 20, 8 << 20, 16 << 20):
        _ops.map_host_multi(xh, "scale", 2.0, out_host=yh, devices=list(range(n_gpus)), chunk_bytes=cb)
        t0 = time.perf_counter()
        for _ in range(5):
            _ops.map_host_multi(xh, "scale", 2.0, out_host=yh, devices=list(range(n_gpus)), chunk_bytes=cb)
        dt = (time.perf_counter() - t0) / 5
        emit(what="map_host_multi_raw", n_gpus=n_gpus, chunk_bytes=cb, ms=dt * 1e3, gbps=2 * n * 4 / dt / 1e9)


if __name__ == "__main__":
    main()

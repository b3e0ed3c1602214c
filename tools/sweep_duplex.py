#!/usr/bin/env python
"""NVLink duplex matrix between GPU 0 (the "root") and GPU 1..N-1: which mechanism moves bytes both ways
at once fastest?  Every pattern moves `nbytes` out of the root and `nbytes` into the root per iteration.

  push/push    root kernel stores into the peers, peer kernels store into the root   (posted writes both ways)
  pull/pull    peer kernels load from the root, root kernel loads from the peers
  rank-driven  every peer runs pull (root->peer) and push (peer->root): fused in one kernel, or as two kernels
  root-driven  the root runs both directions itself
  ce/...       copy engines for one or both directions

Writes JSON lines to gpurun_out/sweep_duplex.jsonl.  With --ncu-one PATTERN runs a single pattern a few times
(for `ncu --section Nvlink*`)."""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from kubetorch_b200.device import lib as L  # noqa: E402
from kubetorch_b200.device import ops  # noqa: E402

OUT = os.path.join(REPO, "gpurun_out", "sweep_duplex.jsonl")
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def emit(f, **kw):
    line = json.dumps(kw)
    print(line, flush=True)
    f.write(line + "\n")
    f.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=224, help="MiB leaving (and entering) the root per iteration")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ncu-one", default=None)
    args = ap.parse_args()
    n_dev = torch.cuda.device_count()
    assert n_dev >= 2
    devs = list(range(n_dev))
    ops.ensure_init(devs)
    peers = devs[1:]
    per = (args.mib << 20) // len(peers) // 4096 * 4096   # bytes per peer per direction
    n = per // 4
    f = open(OUT, "a")
    # root-side source/destination slices, one per peer; peer-side staging in/out
    root_src = [torch.randn(n, device="cuda:0") for _ in peers]
    root_dst = [torch.empty(n, device="cuda:0") for _ in peers]
    peer_in = [torch.empty(n, device=f"cuda:{d}") for d in peers]
    peer_out = [torch.randn(n, device=f"cuda:{d}") for d in peers]
    streams = {d: [torch.cuda.Stream(d), torch.cuda.Stream(d)] for d in devs}

    def kmap(src, dst, dev, variant, stream):
        with torch.cuda.device(dev), torch.cuda.stream(stream):
            ops.map_tensor(src, "identity", out=dst, variant=variant, device=dev, stream=stream)

    def ce(src, dst, dev, stream):
        with torch.cuda.device(dev), torch.cuda.stream(stream):
            dst.copy_(src, non_blocking=True)

    V, T = L.VARIANT_VEC, L.VARIANT_TMA

    def pattern(name):
        """Returns a function enqueuing ONE iteration of the pattern on all devices."""
        def run():
            for i, d in enumerate(peers):
                s_root_a, s_root_b = streams[0]
                s_peer_a, s_peer_b = streams[d]
                if name == "push/push:vec":
                    kmap(root_src[i], peer_in[i], 0, V, s_root_a)
                    kmap(peer_out[i], root_dst[i], d, V, s_peer_a)
                elif name == "push/push:tma":
                    kmap(root_src[i], peer_in[i], 0, T, s_root_a)
                    kmap(peer_out[i], root_dst[i], d, T, s_peer_a)
                elif name == "pull/pull:vec":
                    kmap(root_src[i], peer_in[i], d, V, s_peer_a)
                    kmap(peer_out[i], root_dst[i], 0, V, s_root_a)
                elif name == "rank-driven:2kernels:vec":
                    kmap(root_src[i], peer_in[i], d, V, s_peer_a)
                    kmap(peer_out[i], root_dst[i], d, V, s_peer_b)
                elif name == "rank-driven:2kernels:tma":
                    kmap(root_src[i], peer_in[i], d, T, s_peer_a)
                    kmap(peer_out[i], root_dst[i], d, T, s_peer_b)
                elif name == "rank-driven:fused:vec":
                    kmap(root_src[i], root_dst[i], d, V, s_peer_a)
                elif name == "root-driven:2kernels:vec":
                    kmap(root_src[i], peer_in[i], 0, V, s_root_a)
                    kmap(peer_out[i], root_dst[i], 0, V, s_root_b)
                elif name == "root-push+rank-pull... n/a":
                    pass
                elif name == "ce/ce":
                    ce(root_src[i], peer_in[i], 0, s_root_a)
                    ce(peer_out[i], root_dst[i], d, s_peer_a)
                elif name == "ce-out/push-in":
                    ce(root_src[i], peer_in[i], 0, s_root_a)
                    kmap(peer_out[i], root_dst[i], d, V, s_peer_a)
                elif name == "push-out/ce-in":
                    kmap(root_src[i], peer_in[i], 0, V, s_root_a)
                    ce(peer_out[i], root_dst[i], d, s_peer_a)
                elif name == "pull-out/ce-in":
                    kmap(root_src[i], peer_in[i], d, V, s_peer_a)
                    ce(peer_out[i], root_dst[i], 0, s_root_a)
                elif name == "out-only:push:vec":
                    kmap(root_src[i], peer_in[i], 0, V, s_root_a)
                elif name == "out-only:pull:vec":
                    kmap(root_src[i], peer_in[i], d, V, s_peer_a)
                elif name == "in-only:push:vec":
                    kmap(peer_out[i], root_dst[i], d, V, s_peer_a)
                elif name == "in-only:pull:vec":
                    kmap(peer_out[i], root_dst[i], 0, V, s_root_a)
                else:
                    raise KeyError(name)
        return run

    def sync_all():
        for d in devs:
            torch.cuda.synchronize(d)

    names = ["out-only:push:vec", "out-only:pull:vec", "in-only:push:vec", "in-only:pull:vec",
             "push/push:vec", "push/push:tma", "pull/pull:vec", "rank-driven:2kernels:vec",
             "rank-driven:2kernels:tma", "rank-driven:fused:vec", "root-driven:2kernels:vec", "ce/ce",
             "ce-out/push-in", "push-out/ce-in", "pull-out/ce-in"]
    if args.ncu_one:
        run = pattern(args.ncu_one)
        for _ in range(3):
            run()
        sync_all()
        return
    for name in names:
        run = pattern(name)
        try:
            for _ in range(3):
                run()
            sync_all()
            t0 = time.perf_counter()
            for _ in range(args.iters):
                run()
            sync_all()
            ms = (time.perf_counter() - t0) / args.iters * 1e3
            total = per * len(peers)
            both = not (name.startswith("out-only") or name.startswith("in-only"))
            emit(f, what="duplex", n_dev=n_dev, pattern=name, mib_per_dir=total >> 20, ms=ms,
                 gbps_per_dir=total / ms / 1e6, both_directions=both)
        except Exception as e:  # noqa: BLE001
            emit(f, what="duplex", pattern=name, error=f"{type(e).__name__}: {e}"[:300])
    # parity of the last kernel patterns
    ok = all(torch.equal(peer_in[i].cpu(), root_src[i].cpu()) and torch.equal(root_dst[i].cpu(), peer_out[i].cpu())
             for i in range(len(peers)))
    emit(f, what="duplex_parity", ok=bool(ok))


if __name__ == "__main__":
    main()

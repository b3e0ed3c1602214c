#!/usr/bin/env python
"""C4 policy MLP (256->1024->1024->64, bf16) on one B200: TFLOP/s of ktb_mlp_bf16 vs torch (cuBLAS)."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from kubetorch_b200.device import mlp, ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ops.ensure_init([0])
    g = torch.Generator(device="cuda").manual_seed(0)
    w1 = (torch.randn(1024, 256, device="cuda", generator=g) * 0.02).bfloat16()
    w2 = (torch.randn(1024, 1024, device="cuda", generator=g) * 0.02).bfloat16()
    w3 = (torch.randn(64, 1024, device="cuda", generator=g) * 0.02).bfloat16()
    flop_per_row = 2 * (256 * 1024 + 1024 * 1024 + 1024 * 64)
    # (rows, chunk rows, TMA-store epilogue, CTA-pair kernel, epilogue groups, layer-2+head fused)
    for M, chunk, tst, two_sm, epi, fused in ((262144, 65536, 1, 1, 1, 0), (262144, 65536, 1, 1, 1, 1),
                                              (262144, 75776, 1, 1, 1, 1), (262144, 37888, 1, 1, 1, 1),
                                              (2097152, 65536, 1, 1, 1, 0), (2097152, 75776, 1, 1, 1, 1),
                                              (2097152, 151552, 1, 1, 1, 1)):
        persistent = 1
        ops.set_tuning(7, persistent)
        ops.set_tuning(8, chunk)
        ops.set_tuning(9, epi)
        ops.set_tuning(10, tst)
        ops.set_tuning(11, two_sm)
        ops.set_tuning(18, fused)
        mlp._scratch.clear()
        obs = torch.randn(M, 256, device="cuda", generator=g).bfloat16()
        out = torch.empty(M, 64, dtype=torch.bfloat16, device="cuda")
        ms = timeit(lambda: mlp.mlp_forward(obs, w1, w2, w3, out=out))

        def torch_mlp():
            h = torch.relu(obs @ w1.t())
            h = torch.relu(h @ w2.t())
            return h @ w3.t()

        ms_t = timeit(torch_mlp, iters=5)
        ref = torch_mlp()
        err = (out.float() - ref.float()).abs().max().item()
        print(json.dumps({"what": "mlp_c4", "M": M, "persistent": persistent, "chunk_rows": chunk, "epi_groups": epi, "tma_store": tst, "two_sm": two_sm, "fused_head": fused, "ms": ms, "tflops": flop_per_row * M / ms / 1e9,
                          "torch_ms": ms_t, "torch_tflops": flop_per_row * M / ms_t / 1e9,
                          "max_abs_diff_vs_torch": err,
                          "arg_plus_result_gbps": (M * 256 * 2 + M * 64 * 2) / ms / 1e6}), flush=True)
        del obs, out


if __name__ == "__main__":
    main()

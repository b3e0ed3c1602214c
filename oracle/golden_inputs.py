"""Seeded inputs of the golden cases (shared by oracle/make_golden.py and the tests).

TEST INFRASTRUCTURE.  torch's CPU generator is deterministic for a given seed and call order, so
large inputs (the MLP weights) are regenerated instead of stored; their sha256 is in the fixture.
"""
import hashlib


def tensor_sha256(t) -> str:
    import torch

    return hashlib.sha256(t.contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()


def make_inputs():
    """All tensor inputs, built once (seeded) and stored in the fixture next to the outputs."""
    import torch

    g = torch.Generator().manual_seed(0)
    d_in, d_h, d_out = 256, 1024, 64
    return {
        "f32_1003": torch.arange(1003, dtype=torch.float32) * 0.5 - 100.0,
        "f32_rand_1001": torch.randn(1001, generator=g),
        "u8_1000": torch.randint(0, 256, (1000,), dtype=torch.uint8, generator=g),
        "bf16_777": torch.randn(777, generator=g).bfloat16(),
        "i32_515": torch.randint(-(2**31), 2**31 - 1, (515,), dtype=torch.int32, generator=g),
        "i64_130": torch.randint(-(2**40), 2**40, (130,), dtype=torch.int64, generator=g),
        "f32_3": torch.tensor([1.0, 2.0, 3.0]),
        "f16_515": (torch.randn(515, generator=torch.Generator().manual_seed(16)) * 8).half(),
        "mlp_obs": torch.randn(256, d_in, generator=g).bfloat16(),
        "mlp_w1": (torch.randn(d_h, d_in, generator=g) * 0.02).bfloat16(),
        "mlp_w2": (torch.randn(d_h, d_h, generator=g) * 0.02).bfloat16(),
        "mlp_w3": (torch.randn(d_out, d_h, generator=g) * 0.02).bfloat16(),
    }



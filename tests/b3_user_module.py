"""User module deployed behind the REFERENCE server in tests/test_b3_seam.py: plain kubetorch-style SPMD
functions (they shard by RANK/WORLD_SIZE) plus the one new line that declares the device op."""
import os

import kubetorch_b200 as ktb


def _shard(x):
    r, w = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    chunks = x.chunk(w)
    return chunks[r] if r < len(chunks) else x[:0]


@ktb.mapped("scale", alpha=2.0)
def double(x):
    return _shard(x) * 2


@ktb.mapped("affine", alpha="alpha", beta="beta")
def affine(x, alpha, beta):
    return _shard(x) * alpha + beta


@ktb.mapped("identity", reduce="sum")
def shard_sum(x):
    return int(_shard(x).sum())


class Scaler:
    """kt.cls with a mapped method: one instance per deployment, constructed from KT_INIT_ARGS."""

    def __init__(self, tag="none"):
        self.tag = tag

    @ktb.mapped("scale", alpha=3.0)
    def triple(self, x):
        return _shard(x) * 3

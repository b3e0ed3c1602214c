"""Split a call payload into a small picklable header and its CUDA-tensor leaves.

On the reference's wire every tensor is pickled, base64-encoded and JSON-wrapped once on the client
and decoded once per rank (kt/serving/utils.py:730-749, http_server.py:1768-1822).  On the GPU
rank-process route tensor leaves stay in HBM: they are packed into an arena (ktb_pack), the arena is
replicated to every rank's GPU (ktb_broadcast) and only this header crosses the pipe."""
from __future__ import annotations

from typing import Any, List, Tuple


class TensorRef:
    """Placeholder for the i-th tensor leaf of a payload."""

    __slots__ = ("index", "dtype", "shape")

    def __init__(self, index: int, dtype: str, shape: Tuple[int, ...]):
        self.index, self.dtype, self.shape = index, dtype, tuple(shape)

    def __reduce__(self):
        return (TensorRef, (self.index, self.dtype, self.shape))


def split_tensors(obj: Any, leaves: List, predicate) -> Any:
    """Deep-copy the list/tuple/dict skeleton of `obj`, moving tensors for which predicate(t) holds into
    `leaves` and leaving TensorRef placeholders."""
    import torch

    if isinstance(obj, torch.Tensor) and predicate(obj):
        leaves.append(obj)
        return TensorRef(len(leaves) - 1, str(obj.dtype).replace("torch.", ""), tuple(obj.shape))
    if isinstance(obj, list):
        return [split_tensors(o, leaves, predicate) for o in obj]
    if isinstance(obj, tuple):
        return tuple(split_tensors(o, leaves, predicate) for o in obj)
    if isinstance(obj, dict):
        return {k: split_tensors(v, leaves, predicate) for k, v in obj.items()}
    return obj


def join_tensors(obj: Any, leaves: List) -> Any:
    if isinstance(obj, TensorRef):
        return leaves[obj.index]
    if isinstance(obj, list):
        return [join_tensors(o, leaves) for o in obj]
    if isinstance(obj, tuple):
        return tuple(join_tensors(o, leaves) for o in obj)
    if isinstance(obj, dict):
        return {k: join_tensors(v, leaves) for k, v in obj.items()}
    return obj


def collect_refs(obj: Any, out: List[TensorRef]) -> None:
    if isinstance(obj, TensorRef):
        out.append(obj)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            collect_refs(o, out)
    elif isinstance(obj, dict):
        for v in obj.values():
            collect_refs(v, out)

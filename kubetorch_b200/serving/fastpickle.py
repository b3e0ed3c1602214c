"""Pickle for the coordinator <-> rank-process pipes with a direct route for plain CPU tensors.

The stock reduction of a CPU tensor goes through its storage (`torch.storage._load_from_bytes` → `torch.load` on the
receiving side: ~150 us per tensor, and a view drags its WHOLE storage along — a rank returning `x.chunk(w)[r]` would
send all of x).  On the reference's wire that cost is paid per rank per call (kt/serving/http_server.py:1768-1842,
process_pool.py:178-234).  Here a plain strided CPU tensor is reduced to (dtype, shape, raw bytes): one memcpy each
way, only the bytes the tensor addresses.  Everything else (subclasses, requires_grad, sparse/quantized, CUDA, any
Python object) takes the stock path, so the set of picklable objects is unchanged.  `loads` is plain pickle.loads."""
from __future__ import annotations

import io
import pickle
from typing import Any


def _rebuild_cpu_tensor(dtype_name: str, shape, raw):
    import torch

    dtype = getattr(torch, dtype_name)
    if len(raw) == 0:
        return torch.empty(shape, dtype=dtype)
    buf = raw if isinstance(raw, bytearray) else bytearray(raw)   # torch.frombuffer needs a writable buffer
    return torch.frombuffer(buf, dtype=torch.uint8).view(dtype).reshape(shape)


class _Pickler(pickle.Pickler):
    def reducer_override(self, obj):
        try:
            import torch
        except ImportError:  # pragma: no cover - torch is a dependency of the package
            return NotImplemented
        if type(obj) is torch.Tensor and obj.device.type == "cpu" and obj.layout == torch.strided \
                and not obj.requires_grad and not obj.is_quantized and not obj.is_conj() and not obj.is_neg():
            flat = obj.detach().contiguous().reshape(-1)
            raw = flat.view(torch.uint8).numpy() if flat.numel() else b""
            return _rebuild_cpu_tensor, (str(obj.dtype).replace("torch.", ""), tuple(obj.shape),
                                         pickle.PickleBuffer(raw) if flat.numel() else b"")
        return NotImplemented


def dumps(obj: Any) -> bytes:
    f = io.BytesIO()
    _Pickler(f, protocol=5).dump(obj)
    return f.getvalue()


loads = pickle.loads

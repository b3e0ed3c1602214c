"""bf16 MLP policy (BASELINE config C4) on tcgen05 tensor cores — tensor-facing wrapper of
ktb_mlp_bf16 and its scatter→exec→gather form."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import lib as L
from . import ops

_scratch = {}
_weight_cache = {}


def _scratch_for(dev: int, M: int, d_hidden: int) -> torch.Tensor:
    nbytes = L.load().ktb_mlp_scratch_bytes(M, d_hidden)
    buf = _scratch.get(dev)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
        _scratch[dev] = buf
    return buf


def mlp_forward(obs: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor, w3: torch.Tensor,
                out: Optional[torch.Tensor] = None, device: Optional[int] = None,
                stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """logits[M, d_out] = W3·relu(W2·relu(W1·obsᵀ)); bf16 storage, fp32 accumulation in TMEM.
    `obs` / `out` may be peer-mapped (pull the observations / push the logits over NVLink)."""
    for name, t in (("obs", obs), ("w1", w1), ("w2", w2), ("w3", w3)):
        if t.dtype != torch.bfloat16 or not t.is_cuda or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous CUDA bfloat16 tensor")
    dev = w1.device.index if device is None else int(device)
    ops.ensure_init({dev, obs.device.index})
    M, d_in = obs.shape
    d_hidden, d_out = w1.shape[0], w3.shape[0]
    if w1.shape != (d_hidden, d_in) or w2.shape != (d_hidden, d_hidden) or w3.shape != (d_out, d_hidden):
        raise ValueError("weight shapes must be W1[d_h,d_in], W2[d_h,d_h], W3[d_out,d_h] (nn.Linear layout)")
    if out is None:
        out = torch.empty(M, d_out, dtype=torch.bfloat16, device=f"cuda:{dev}")
    else:
        ops.ensure_init({out.device.index})
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    L.call("ktb_mlp_bf16", dev, obs.data_ptr(), M, d_in, d_hidden, d_out, w1.data_ptr(), w2.data_ptr(), w3.data_ptr(),
           out.data_ptr(), _scratch_for(dev, M, d_hidden).data_ptr(), int(s.cuda_stream))
    return out


def _weights_on(dev: int, ws: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Weights are shared by all ranks: broadcast once per (weights, device) and kept resident
    (excluded from per-call bytes)."""
    out = []
    for w in ws:
        if w.device.index == dev:
            out.append(w)
            continue
        key = (w.data_ptr(), w._version, dev)
        c = _weight_cache.get(key)
        if c is None:
            c = torch.empty_like(w, device=f"cuda:{dev}")
            ops.broadcast(w, [c])
            torch.cuda.synchronize(w.device)
            _weight_cache[key] = c
        out.append(c)
    return out


def mlp_scatter_gather(obs_root: torch.Tensor, w1, w2, w3, devices: Sequence[int],
                       out_root: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """Rank r runs the MLP on `obs.chunk(world)[r]`: its first GEMM's TMA loads read the rows straight
    from the root GPU (scatter) and its last epilogue stores the logits straight into the root's
    result buffer (gather). Returns rank-ordered views of the root result."""
    root = obs_root.device.index
    devs = [int(d) for d in devices]
    if devs[0] != root:
        raise ValueError("obs must live on the root GPU (devices[0])")
    ops.ensure_init(set(devs))
    M = obs_root.shape[0]
    d_out = w3.shape[0]
    if out_root is None:
        out_root = torch.empty(M, d_out, dtype=torch.bfloat16, device=obs_root.device)
    root_stream = torch.cuda.current_stream(root)
    ready = torch.cuda.Event()
    ready.record(root_stream)
    views, done = [], []
    for r, dev in enumerate(devs):
        b, e = ops.shard_bounds(M, len(devs), r)
        views.append(out_root[b:e])
        if e == b:
            continue
        ws = _weights_on(dev, (w1, w2, w3))
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev)
            if dev != root:
                st.wait_event(ready)
            mlp_forward(obs_root[b:e], ws[0], ws[1], ws[2], out=out_root[b:e], device=dev, stream=st)
            if dev != root:
                ev = torch.cuda.Event()
                ev.record(st)
                done.append(ev)
    for ev in done:
        root_stream.wait_event(ev)
    return views

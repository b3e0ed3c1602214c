"""The B200 dispatch backend: distribution_type "b200" behind the reference's supervisor seam
(kt/serving/supervisor_factory.py:11-58 — "new distribution_type values are added here").

One controller process drives N local B200s.  For a @kt.mapped callable a remote call becomes:

    device-resident arg (CUDA tensor on the root GPU)
        ktb_scatter_map_gather: rank r's kernel pulls `x.chunk(N)[r]` from the root's HBM over
        NVLink/NVSwitch, applies the op, pushes the result into the root's result buffer —
        scatter, exec and gather are ONE kernel per rank; the caller gets N views, rank-ordered.
    host-resident arg (CPU tensor; the reference's client lives outside the GPU)
        each rank's shard goes host → its own GPU → host over that GPU's own PCIe link
        (ktb_map_host: chunked H2D / kernel / D2H on three streams), ranks in parallel.
    reduce="sum"
        ktb_scatter_map_reduce: per-rank warp-shuffle reduction, scalar peer-stored to the root.

Result shape follows the reference (spmd_supervisor.py:547-570, execution_supervisor.py:141):
a rank-ordered list with one entry per rank when the compute is distributed, the bare value
otherwise.  There is NO CPU fallback here: a missing library or GPU raises.
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional

from ..mapped import ELEMENTWISE_OPS, mapped_spec
from .codec import HTTPException, check_allowed
from .supervisors import check_callable_name
from .process_worker import instantiate, load_callable, resolve_method


class B200Supervisor:
    def __init__(self, pointers=None, init_args=None, name: str = None, devices: Optional[List[int]] = None,
                 num_proc=None, distributed: bool = True, allowed_serialization: str = "json,pickle",
                 host_chunk_bytes: int = 16 << 20, variant: int = 0, callable_obj=None, transfer: str = "auto",
                 host_mode: str = "multi", placement: str = "ranks", **extra):
        self.pointers, self.init_args, self.name = pointers, init_args, name
        self.callable_obj = callable_obj
        self.devices = list(devices) if devices is not None else None
        self.num_proc = num_proc
        self.distributed = distributed
        self.allowed_serialization = allowed_serialization
        self.host_chunk_bytes = host_chunk_bytes
        self.variant = variant
        self._callable = None
        self._host_pool: Optional[ThreadPoolExecutor] = None
        self._pin_cache = {}
        self._lock = threading.Lock()
        self._host_lock = threading.Lock()
        self._push = None
        if transfer not in ("auto", "pull", "push"):
            raise ValueError("transfer must be 'auto', 'pull' or 'push'")
        self.transfer = transfer
        self.host_mode = host_mode  # "multi": one C call drives all GPUs; "threads": one host thread per rank
        # where device-resident element-wise shards execute: "ranks" = rank r on GPU r (the sharded path the
        # benchmarks measure); "root" = every rank's shard on the root GPU (same results; for HBM-bound ops a
        # root-resident arg is served faster by the root's own HBM than through its NVLink port, profiles §2)
        if placement not in ("ranks", "root"):
            raise ValueError("placement must be 'ranks' or 'root'")
        self.placement = placement
        self.config_hash = hash(("b200", tuple(self.devices or ()), num_proc, distributed))

    # ---- lifecycle ------------------------------------------------------------------------------------
    def setup(self):
        import torch

        from ..device import lib as L
        from ..device import ops

        L.load()  # fail loudly if libktb200.so is missing
        ops.require_cuda()
        if self.devices is None:
            n = self.num_proc
            if n in (None, "auto"):
                n = torch.cuda.device_count()
            self.devices = list(range(int(n)))
        if self.devices and max(self.devices) >= torch.cuda.device_count():
            raise RuntimeError(
                f"kt.Compute asked for GPU index {max(self.devices)} ({len(self.devices)} ranks) but only "
                f"{torch.cuda.device_count()} GPUs are visible"
            )
        ops.ensure_init(self.devices)
        self._callable = instantiate(self.callable_obj, self.init_args) if self.callable_obj is not None \
            else load_callable(self.pointers, self.init_args)
        self._host_pool = ThreadPoolExecutor(max_workers=len(self.devices), thread_name_prefix="ktb-host")

    def cleanup(self):
        self._callable = None
        if self._host_pool is not None:
            self._host_pool.shutdown(wait=False)
            self._host_pool = None
        self._pin_cache.clear()
        self._push = None

    @property
    def world_size(self) -> int:
        return len(self.devices)

    # ---- the call ---------------------------------------------------------------------------------------
    def call(self, request, cls_or_fn_name, method_name=None, params=None, distributed_subcall=False):
        import torch

        serialization = request.headers.get("X-Serialization", "json")
        if self._callable is None:
            raise HTTPException(503, "Server is loading the callable. Please retry in a moment.")
        check_callable_name(cls_or_fn_name, self.name)
        check_allowed(serialization, self.allowed_serialization)
        params = params or {}
        method = resolve_method(self._callable, cls_or_fn_name, method_name)
        spec = mapped_spec(method)
        if spec is None:
            raise TypeError(
                f"'{cls_or_fn_name}' is not a @kt.mapped callable: the 'b200' backend executes registered device ops "
                "only. Use .distribute('spmd'|'pytorch', num_proc=N) to run arbitrary Python on N local ranks."
            )
        x, alpha, beta, bound = spec.bind(method, params.get("args", []), params.get("kwargs", {}))
        if spec.op == "mlp":
            return self._call_mlp(spec, bound)
        if spec.op not in ELEMENTWISE_OPS:
            raise TypeError(f"unsupported mapped op '{spec.op}'")
        if not isinstance(x, torch.Tensor):
            raise TypeError(f"mapped callable '{cls_or_fn_name}' expects a torch.Tensor argument, got {type(x).__name__}")
        if spec.reduce == "sum":
            out = self._reduce(x, spec.op, alpha, beta)
        elif x.is_cuda:
            out = self._device_map(x, spec.op, alpha, beta)
        else:
            out = self._host_map(x, spec.op, alpha, beta)
        return out if self.distributed else out[0] if len(out) == 1 else out

    def _shard_views(self, flat_like, x):
        """Rank-ordered views of `flat_like` (same layout as x) following x.chunk(world) on dim 0."""
        from ..device import ops

        rows = x.shape[0] if x.dim() > 0 else 1
        y = flat_like.view(x.shape) if x.dim() > 0 else flat_like.view(1)
        if self.world_size == 1:
            return [y]
        views = []
        for r in range(self.world_size):
            b, e = ops.shard_bounds(rows, self.world_size, r)
            views.append(y[b:e])
        return views

    def _device_map(self, x, op, alpha, beta):
        import torch

        from ..device import ops

        root = self.devices[0]
        if x.device.index != root:
            raise ValueError(f"device-resident args must live on the root GPU cuda:{root}, got {x.device}")
        if x.dim() == 0:
            x = x.reshape(1)
        x = x.contiguous()
        out = torch.empty_like(x)  # same device as x: the root GPU
        if self.placement == "root":
            ops.scatter_map_gather(x, op, alpha, beta, devices=[root] * len(self.devices), out_root=out,
                                   variant=self.variant)
            return self._shard_views(out, x)
        distinct = len(set(self.devices)) == len(self.devices) and len(self.devices) > 1
        nbytes = x.numel() * x.element_size()
        # measured crossover (profiles/r1_summary.md §2): the flag pipeline wins from 256 MiB at N >= 4 and from 1 GiB
        # at N = 2; below that the single fused kernel per rank has less fixed cost
        auto_push = distinct and ((len(self.devices) >= 4 and nbytes >= (256 << 20)) or nbytes >= (1 << 30))
        if self.transfer == "push" or (self.transfer == "auto" and auto_push):
            # push/push flag pipeline: both NVLink directions carry posted writes (see ktb_push.cu)
            with self._lock:
                rows = x.shape[0]
                shard_bytes = ops.shard_bounds(rows, self.world_size, 0)[1] * ops.row_elems(x) * x.element_size()
                if self._push is None or self._push.stride < shard_bytes:
                    self._push = ops.PushSession(self.devices, shard_bytes)
                self._push.call(x, out, op, alpha, beta)
        else:
            ops.scatter_map_gather(x, op, alpha, beta, devices=self.devices, out_root=out, variant=self.variant)
        return self._shard_views(out, x)

    def _pinned(self, key, like):
        import torch

        buf = self._pin_cache.get(key)
        if buf is None or buf.numel() != like.numel() or buf.dtype != like.dtype:
            buf = torch.empty(like.numel(), dtype=like.dtype).pin_memory()
            self._pin_cache[key] = buf
        return buf.view(like.shape)

    def _host_map(self, x, op, alpha, beta):
        from ..device import ops

        if x.dim() == 0:
            x = x.reshape(1)
        x = x.contiguous()
        with self._host_lock:  # the pinned staging tensors are per deployment
            return self._host_map_locked(x, op, alpha, beta)

    def _host_map_locked(self, x, op, alpha, beta):
        from ..device import ops

        if not x.is_pinned():
            staged = self._pinned("in", x)
            staged.copy_(x)  # page-locking copy: the caller handed us pageable memory
            x = staged
        # results are FRESH tensors every call (the reference returns new objects); torch's caching host allocator
        # makes repeated page-locked allocations of the same size cheap
        import torch

        out = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
        x_shards = self._shard_views(x, x)
        o_shards = self._shard_views(out, x)
        if self.host_mode == "multi" and len(set(self.devices)) == len(self.devices):
            # one C call drives every GPU's copy/exec/copy pipeline (no per-rank Python threads)
            ops.map_host_multi(x, op, alpha, beta, out_host=out, devices=self.devices)
            return o_shards

        def run(rank):
            if x_shards[rank].numel():
                ops.map_host(x_shards[rank], op, alpha, beta, out_host=o_shards[rank], device=self.devices[rank],
                             chunk_bytes=self.host_chunk_bytes)

        if self.world_size == 1:
            run(0)
        else:
            list(self._host_pool.map(run, range(self.world_size)))
        return o_shards

    def _reduce(self, x, op, alpha, beta):
        import torch

        from ..device import ops

        root = self.devices[0]
        if not x.is_cuda:
            x = x.to(f"cuda:{root}", non_blocking=True)
        elif x.device.index != root:
            raise ValueError(f"device-resident args must live on the root GPU cuda:{root}, got {x.device}")
        if x.dim() == 0:
            x = x.reshape(1)
        x = x.contiguous()
        _, partials = ops.scatter_map_reduce(x, op, alpha, beta, devices=self.devices)
        return partials.tolist()  # per-rank Python scalars, as the reference's ranks return

    def _call_mlp(self, spec, bound):
        from ..device import mlp

        names = list(bound)
        obs, w1, w2, w3 = (bound[n] for n in names[:4])
        out = mlp.mlp_scatter_gather(obs, w1, w2, w3, devices=self.devices)
        return out if self.distributed else out[0] if len(out) == 1 else out

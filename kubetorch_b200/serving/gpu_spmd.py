"""SPMD rank processes on GPUs with HBM arenas as the wire — the reference's broadcast semantics
(every rank sees the full args, spmd_supervisor.py:341) for ARBITRARY Python callables, without the
pickle → base64 → JSON → queue-per-rank path for tensors:

    coordinator  ktb_pack      tensor leaves of (args, kwargs) → arg arena on the root GPU
                 ktb_broadcast one kernel: root arena → every other rank's arg arena (NVLink peer stores)
                 pipe          small pickled header (tensor placeholders + offsets) to every rank
    rank r       zero-copy views of its local arg arena → user callable → ktb_pack result leaves into
                 its result arena → small header back
    coordinator  ktb_unpack    result arena of rank r → fresh tensors on the root GPU (peer stores)

The arenas are owned by the coordinator (one per rank GPU, CUDA-IPC exported to the rank process) and
allocated lazily on the first call that carries CUDA tensors, so launcher-only jobs (DDP) never touch them.
"""
from __future__ import annotations

import os
import pickle
import threading
from concurrent.futures import FIRST_EXCEPTION, wait
from typing import Dict, List, Optional

from .codec import HTTPException, check_allowed
from . import fastpickle
from .process_pool import ProcessPool
from .supervisors import SPMDSupervisor, check_callable_name, select_worker_nodes
from .tensor_wire import collect_refs, join_tensors, split_tensors


class GpuSPMDSupervisor(SPMDSupervisor):
    def __init__(self, *args, devices: Optional[List[int]] = None, arena_bytes: int = 64 << 20, **kwargs):
        super().__init__(*args, **kwargs)
        self.devices = list(devices) if devices else None
        self.initial_arena_bytes = int(arena_bytes)
        self.arg_arenas: List = []
        self.res_arenas: List = []
        self._pending_updates: Dict[int, dict] = {}
        self._retired: List = []          # superseded arenas: freed only after the ranks closed their IPC mappings
        self._call_lock = threading.Lock()

    # ---- lifecycle ---------------------------------------------------------------------------------------
    def setup(self):
        import torch

        if self.pool is not None:
            self.cleanup()
        n_dev = torch.cuda.device_count() if torch.cuda.is_available() else 0
        if self.devices is None:
            self.devices = [r % max(n_dev, 1) for r in range(self.world_size)]
        pod_names = [f"{self.name}-{r // self.num_proc}" for r in range(self.world_size)]
        gpu_cfgs = [{"device": d} for d in self.devices] if n_dev else None
        if n_dev and "KTB_STORE_DIR" not in self.env_vars:
            import tempfile

            base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
            self._store_dir = tempfile.mkdtemp(prefix=f"ktb200_store_{os.getpid()}_", dir=base)
            self.env_vars["KTB_STORE_DIR"] = self._store_dir   # kt.put / kt.get between the rank processes
        self.pool = ProcessPool(
            self.world_size, self.pointers, self.init_args, self.name, max_threads_per_proc=self.max_threads_per_proc,
            base_env=self.env_vars, allowed_serialization=self.allowed_serialization, pod_names=pod_names,
            gpu_cfgs=gpu_cfgs,
        )

    def cleanup(self):
        super().cleanup()
        for a in self.arg_arenas + self.res_arenas:
            try:
                a.free()
            except Exception:  # noqa: BLE001
                pass
        for _, a in self._retired:
            try:
                a.free()
            except Exception:  # noqa: BLE001
                pass
        self.arg_arenas, self.res_arenas, self._retired = [], [], []
        self._pending_updates = {}
        if getattr(self, "_store_dir", None):
            import shutil

            shutil.rmtree(self._store_dir, ignore_errors=True)
            self.env_vars.pop("KTB_STORE_DIR", None)
            self._store_dir = None

    # ---- arenas --------------------------------------------------------------------------------------------
    def _ensure_arenas(self, arg_bytes: int):
        from ..device import ops

        ops.ensure_init(set(self.devices))
        if not self.arg_arenas:
            cap = max(self.initial_arena_bytes, _round_up(arg_bytes))
            self.arg_arenas = [ops.Arena(d, cap) for d in self.devices]
            self.res_arenas = [ops.Arena(d, cap) for d in self.devices]
            for r in range(self.world_size):
                self._pending_updates[r] = {
                    "arg_handle": self.arg_arenas[r].export(), "arg_bytes": cap,
                    "res_handle": self.res_arenas[r].export(), "res_bytes": cap,
                }
        elif arg_bytes > self.arg_arenas[0].nbytes:
            cap = _round_up(arg_bytes * 2)
            for r in range(self.world_size):
                self._grow(r, "arg", cap)

    def _grow(self, r: int, which: str, cap: int):
        import torch

        from ..device import ops

        arenas = self.arg_arenas if which == "arg" else self.res_arenas
        torch.cuda.synchronize(self.devices[r])
        old = arenas[r]
        arenas[r] = ops.Arena(self.devices[r], cap)
        # freeing exported memory while the rank still maps it is undefined behaviour (CUDA IPC): the rank closes the
        # old mapping when it receives the update; the free happens after that call's reply (_free_retired)
        self._retired.append((r, old))
        upd = self._pending_updates.setdefault(r, {})
        upd[f"{which}_handle"], upd[f"{which}_bytes"] = arenas[r].export(), cap

    def _free_retired(self, replied_ranks):
        """Free superseded arenas whose rank has (a) been sent the new handle and (b) replied since."""
        keep = []
        for r, a in self._retired:
            if r in replied_ranks and r not in self._pending_updates:
                a.free()
            else:
                keep.append((r, a))
        self._retired = keep

    # ---- the call ---------------------------------------------------------------------------------------------
    def call(self, request, cls_or_fn_name, method_name=None, params=None, distributed_subcall=False):
        import torch

        serialization = request.headers.get("X-Serialization", "json")
        params = params or {}
        if self.pool is None:
            raise HTTPException(503, "Server is loading the callable. Please retry in a moment.")
        check_callable_name(cls_or_fn_name, self.name)
        check_allowed(serialization, self.allowed_serialization)
        nodes = select_worker_nodes(params.get("workers"), self.worker_ips, self.worker_ips[0])
        if params.get("restart_procs", False):
            self.cleanup()
            self.setup()
        ranks = [n * self.num_proc + l for n in nodes for l in range(self.num_proc)]
        args, kwargs = params.get("args", []), params.get("kwargs", {})
        root = self.devices[0]
        leaves: List = []
        use_arenas = torch.cuda.is_available() and serialization != "json"
        # only dtypes the rank-side arena view supports travel through HBM; the rest (complex, float8, uint16/32/64)
        # take the pickle path like any other Python object
        skeleton = split_tensors((args, kwargs), leaves, lambda t: t.is_cuda and str(t.dtype) in _ARENA_DTYPES) \
            if use_arenas else (args, kwargs)
        # The arenas carry one call at a time, so a call that moves CUDA tensors through them is serialised.  Calls
        # WITHOUT tensor args run concurrently (the reference's concurrency model: thread callers and async callables
        # overlap, kt/serving/design.md:67-85); their CUDA results, if any, travel pickled instead of through the
        # result arena.
        import contextlib

        locked = bool(leaves)
        with (self._call_lock if locked else contextlib.nullcontext()):
            extras: Dict[int, dict] = {}
            if leaves:
                from ..device import ops

                leaves = [t.contiguous() if t.device.index == root else t.to(f"cuda:{root}").contiguous() for t in leaves]
                nbytes = [t.numel() * t.element_size() for t in leaves]
                offsets, total = ops.pack_layout(nbytes)
                self._ensure_arenas(total)
                with torch.cuda.device(root):
                    root_arena = self.arg_arenas[0].tensor(torch.uint8)
                    ops.pack(leaves, arena=root_arena)                       # tensor leaves → root arena
                    dsts = []
                    seen = {self.arg_arenas[0].ptr}
                    for r in ranks:
                        a = self.arg_arenas[r]
                        if a.ptr not in seen:
                            seen.add(a.ptr)
                            dsts.append(a.tensor(torch.uint8)[:max(total, 1)])
                    if dsts and total:
                        ops.broadcast(root_arena[:total], dsts)                   # one read, N-1 peer stores
                    torch.cuda.current_stream(root).synchronize()                 # data is in every rank's HBM
                for r in ranks:
                    extras[r] = {"arg_offsets": offsets, "res_arena": True}
                for r in ranks:
                    if r in self._pending_updates:
                        extras.setdefault(r, {})["arena_update"] = self._pending_updates.pop(r)
            payload = fastpickle.dumps(skeleton)   # CPU tensor leaves as raw bytes; CUDA leaves were moved to the arena
            envs = self.rank_envs()
            futures = self.pool.call_all(payload, method_name, envs, serialization, ranks=ranks, extras=extras)
            done, _ = wait(futures, return_when=FIRST_EXCEPTION)
            for f in futures:
                if f in done and f.exception() is not None:
                    raise f.exception()
            results = []
            touched = set()
            for r, f in zip(ranks, futures):
                msg = f.result()
                value = pickle.loads(msg["result"]) if isinstance(msg, dict) else pickle.loads(msg)
                if isinstance(msg, dict) and "res_offsets" in msg:
                    value = self._gather_result(r, value, msg["res_offsets"])
                    touched.add(self.devices[r])
                if isinstance(msg, dict) and msg.get("res_needed"):
                    self._grow(r, "res", _round_up(msg["res_needed"] * 2))
                results.append(value)
            for dev in touched:  # one sync per GPU after ALL gathers are enqueued (they run concurrently)
                torch.cuda.current_stream(dev).synchronize()
            if locked and self._retired:
                self._free_retired(set(ranks))
        return results

    def _gather_result(self, r: int, skeleton, offsets):
        """Result leaves of rank r: result arena on GPU r → fresh tensors on the root GPU (one segmented
        kernel on GPU r storing over NVLink)."""
        import torch

        from ..device import ops

        refs: List = []
        collect_refs(skeleton, refs)
        refs.sort(key=lambda x: x.index)
        root, dev = self.devices[0], self.devices[r]
        with torch.cuda.device(root):
            outs = [torch.empty(ref.shape, dtype=getattr(torch, ref.dtype), device=f"cuda:{root}") for ref in refs]
        with torch.cuda.device(dev):
            arena = self.res_arenas[r].tensor(torch.uint8)
            ops.unpack(arena, offsets, outs)
        return join_tensors(skeleton, outs)


_ARENA_DTYPES = {f"torch.{n}" for n in ("uint8", "int8", "float32", "float64", "int32", "int64", "float16", "bfloat16",
                                           "bool", "int16")}


def _round_up(n: int, align: int = 1 << 20) -> int:
    return max(align, (int(n) + align - 1) // align * align)

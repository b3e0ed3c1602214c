"""Rank worker process: loads the user callable once, then serves calls.

Role of kt/serving/process_worker.py:15-272 (ProcessWorker) on the local route, re-designed:
  * requests arrive on a duplex multiprocessing Connection as raw bytes — ONE pickle of the call
    payload made by the coordinator is shared by all ranks (the reference re-pickles the whole
    base64 body once per rank through mp.Queue, process_pool.py:149-161, its largest cost);
  * no 10 ms queue poll (process_worker.py:198-202): the worker blocks in recv_bytes();
  * sync callables run on a thread pool (max_threads, default 10 under SPMD like
    execution_supervisor.py:39), async callables on one asyncio loop, so concurrent calls overlap
    exactly as in the reference (kt/serving/design.md:67-85);
  * per-request env update with the distributed contract (process_worker.py:75-102,128-129).
"""
from __future__ import annotations

import asyncio
import importlib
import inspect
import json
import os
import pickle
import sys
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, Optional

from ..exceptions import SerializationError
from .codec import check_allowed, package_exception
from . import fastpickle

SHUTDOWN = b"__KTB_SHUTDOWN__"


def load_callable(pointers, init_args: Optional[dict]):
    """Import (root_path, module_name, name) — the reference's pointer triple — and instantiate classes."""
    root_path, module_name, name = pointers
    if root_path and root_path not in sys.path:
        sys.path.insert(0, root_path)
    module = importlib.import_module(module_name)
    obj = module
    for part in name.split("."):
        obj = getattr(obj, part)
    return instantiate(obj, init_args)


def load_callable_from_env():
    """The reference server's loader (kt/serving/http_server.py:1040-1101): KT_FILE_PATH on sys.path, import
    KT_MODULE_NAME, take KT_CLS_OR_FN_NAME, unwrap deploy decorators, instantiate classes from KT_INIT_ARGS.
    Used when a supervisor is built the way `load_supervisor` builds it — from KT_DISTRIBUTED_CONFIG alone."""
    from .codec import HTTPException

    try:
        name, module_name = os.environ["KT_CLS_OR_FN_NAME"], os.environ["KT_MODULE_NAME"]
    except KeyError as e:
        raise RuntimeError(f"{e.args[0]} is not set: no callable metadata to load (deploy with .to() or set KT_* env)")
    root = os.environ.get("KT_FILE_PATH")
    if root:
        root = os.path.abspath(os.path.expanduser(root))
        if root not in sys.path:
            sys.path.insert(0, root)
    module = importlib.import_module(module_name)
    try:
        obj = getattr(module, name)
    except AttributeError as e:
        raise HTTPException(404, f"Callable '{name}' not found in module '{module_name}'") from e
    if getattr(obj, "_kt_partial_module", False) and hasattr(obj, "__wrapped__"):
        obj = obj.__wrapped__
    init_args = None
    raw = os.environ.get("KT_INIT_ARGS", "null")
    if raw not in ("null", "None", ""):
        init_args = json.loads(raw)
    return instantiate(obj, init_args)


def instantiate(obj, init_args: Optional[dict]):
    """Classes become one instance per rank built from init_args (http_server.py:1088-1101)."""
    if inspect.isclass(obj):
        obj = obj(**dict(init_args or {}))
    return obj


def resolve_method(callable_obj, cls_or_fn_name: str, method_name: Optional[str]):
    from .codec import HTTPException

    if method_name:
        if not hasattr(callable_obj, method_name):
            raise HTTPException(404, f"Method '{method_name}' not found in class '{cls_or_fn_name}'")
        return getattr(callable_obj, method_name)
    return callable_obj


def validate_result(result: Any, serialization: str):
    if serialization == "json":
        try:
            json.dumps(result)
        except (TypeError, ValueError) as e:
            raise SerializationError(f"Result could not be serialized to JSON: {e}")
    return result


class _GpuArenas:
    """Rank-side view of the coordinator-owned HBM arenas (CUDA IPC): tensor args arrive as zero-copy views
    of the arg arena, tensor results leave through the result arena (ktb_pack)."""

    def __init__(self, cfg: dict):
        import torch

        from ..device import ops

        self.device = int(cfg["device"])
        torch.cuda.set_device(self.device)
        ops.ensure_init([self.device])
        self.ops = ops
        self.torch = torch
        self.arg_ptr = self.res_ptr = 0
        self.arg_bytes = self.res_bytes = 0
        self.update(cfg)

    def update(self, cfg: dict):
        """Open the new arena handle(s); the superseded mapping is CLOSED first (the coordinator defers the
        cudaFree of the old arena until this rank has replied, see GpuSPMDSupervisor._grow)."""
        if cfg.get("arg_handle") is not None:
            if self.arg_ptr:
                self.torch.cuda.synchronize(self.device)
                self.ops.ipc_close(self.device, self.arg_ptr)
            self.arg_ptr, self.arg_bytes = self.ops.ipc_open(self.device, cfg["arg_handle"]), int(cfg["arg_bytes"])
        if cfg.get("res_handle") is not None:
            if self.res_ptr:
                self.torch.cuda.synchronize(self.device)
                self.ops.ipc_close(self.device, self.res_ptr)
            self.res_ptr, self.res_bytes = self.ops.ipc_open(self.device, cfg["res_handle"]), int(cfg["res_bytes"])

    def _view(self, ptr: int, dtype_name: str, shape, offset: int):
        torch = self.torch
        dtype = getattr(torch, dtype_name)
        numel = 1
        for d in shape:
            numel *= d
        typestr = {"uint8": "|u1", "int8": "|i1", "float32": "<f4", "float64": "<f8", "int32": "<i4", "int64": "<i8",
                   "float16": "<f2", "bfloat16": "<u2", "bool": "|u1", "int16": "<i2"}[dtype_name]
        holder = type("_CAI", (), {})()
        holder.__cuda_array_interface__ = {"shape": (max(numel, 1),), "typestr": typestr,
                                           "data": (ptr + offset, False), "version": 3}
        t = torch.as_tensor(holder, device=f"cuda:{self.device}")[:numel]
        if dtype in (torch.bfloat16, torch.bool):
            t = t.view(dtype)
        return t.reshape(shape)

    def arg_views(self, refs, offsets):
        return [self._view(self.arg_ptr, r.dtype, r.shape, off) for r, off in zip(refs, offsets)]

    def arg_tensors(self, refs, offsets):
        """OWNED copies of the arg leaves (one ktb_unpack launch): the arena is overwritten by the next call's
        pack + broadcast, and the reference hands every call freshly deserialised tensors — a callable may keep
        them (`self.weights = state_dict`)."""
        torch = self.torch
        outs = [torch.empty(tuple(r.shape), dtype=getattr(torch, r.dtype), device=f"cuda:{self.device}") for r in refs]
        if outs:
            arena = self._view(self.arg_ptr, "uint8", (self.arg_bytes,), 0)
            self.ops.unpack(arena, list(offsets), outs)
        return outs

    def pack_results(self, leaves):
        """Pack result leaves into the result arena. Returns offsets, or None if the arena is too small."""
        torch, ops = self.torch, self.ops
        leaves = [t.contiguous() for t in leaves]
        nbytes = [t.numel() * t.element_size() for t in leaves]
        offsets, total = ops.pack_layout(nbytes)
        if total > self.res_bytes:
            return None, total
        arena = self._view(self.res_ptr, "uint8", (self.res_bytes,), 0)
        ops.pack(leaves, arena=arena)
        torch.cuda.current_stream(self.device).synchronize()  # results are in HBM before the reply leaves
        return offsets, total


def worker_main(conn, local_rank: int, pointers, init_args, name: str, max_threads: int, base_env: Dict[str, str],
                allowed_serialization: str, gpu_cfg: Optional[dict] = None):
    os.environ["LOCAL_RANK"] = str(local_rank)  # also set at construction in the reference (process_worker.py:33)
    os.environ.update(base_env)
    send_lock = threading.Lock()
    arena_lock = threading.Lock()
    state = {"arenas": None}

    def reply(msg: dict):
        data = pickle.dumps(msg, protocol=5)
        with send_lock:
            conn.send_bytes(data)

    try:
        callable_obj = load_callable(pointers, init_args)
    except BaseException as e:  # noqa: BLE001
        reply({"id": -1, "ok": False, "envelope": package_exception(e)})
        return
    reply({"id": -1, "ok": True, "result": None})

    executor = ThreadPoolExecutor(max_workers=max_threads, thread_name_prefix=f"ktb-rank{local_rank}")
    loop = asyncio.new_event_loop()
    loop_thread = threading.Thread(target=loop.run_forever, name="ktb-asyncio", daemon=True)
    loop_thread.start()

    def run_request(req: dict):
        try:
            os.environ.update(req.get("env") or {})
            check_allowed(req["serialization"], allowed_serialization)
            args, kwargs = pickle.loads(req["payload"])
            if gpu_cfg is not None and req.get("arena_update"):
                with arena_lock:
                    if state["arenas"] is None:  # CUDA is touched by the framework only when tensors travel
                        state["arenas"] = _GpuArenas({**gpu_cfg, **req["arena_update"]})
                    else:
                        state["arenas"].update(req["arena_update"])
            arenas = state["arenas"]
            if arenas is not None:
                from .tensor_wire import collect_refs, join_tensors

                refs = []
                collect_refs((args, kwargs), refs)
                if refs:
                    refs.sort(key=lambda r: r.index)
                    owned = arenas.arg_tensors(refs, req["arg_offsets"])
                    args, kwargs = join_tensors((args, kwargs), owned)
            method = resolve_method(callable_obj, name, req.get("method"))
            if inspect.iscoroutinefunction(method):
                result = asyncio.run_coroutine_threadsafe(method(*args, **kwargs), loop).result()
            else:
                result = method(*args, **kwargs)
                if inspect.isawaitable(result):
                    async def _await(x):
                        return await x
                    result = asyncio.run_coroutine_threadsafe(_await(result), loop).result()
            extra = {}
            if arenas is not None and req.get("res_arena") and req["serialization"] != "json":
                from .tensor_wire import split_tensors

                leaves = []
                skeleton = split_tensors(result, leaves, lambda t: t.is_cuda and t.device.index == arenas.device)
                if leaves:
                    offsets, total = arenas.pack_results(leaves)
                    if offsets is not None:
                        result, extra = skeleton, {"res_offsets": offsets, "res_bytes": total}
                    else:
                        extra = {"res_needed": total}  # arena too small: this reply travels pickled, next one fits
            validate_result(result, req["serialization"])
            try:
                payload = fastpickle.dumps(result)   # plain CPU tensors as raw bytes (see fastpickle.py)
            except Exception as e:  # noqa: BLE001
                raise SerializationError(f"Result could not be serialized with pickle: {e}")
            reply({"id": req["id"], "ok": True, "result": payload, **extra})
        except BaseException as e:  # noqa: BLE001
            reply({"id": req["id"], "ok": False, "envelope": package_exception(e)})

    while True:
        try:
            data = conn.recv_bytes()
        except (EOFError, OSError):
            break
        if data == SHUTDOWN:
            break
        try:
            req = pickle.loads(data)
        except Exception as e:  # noqa: BLE001 - a damaged frame is reported, it does not take the rank down
            reply({"id": -2, "ok": False, "envelope": package_exception(e)})
            continue
        executor.submit(run_request, req)

    executor.shutdown(wait=False, cancel_futures=True)
    loop.call_soon_threadsafe(loop.stop)
    # framework cleanup on reload (kt/serving/spmd/pytorch_process.py:8-16)
    try:
        if "torch.distributed" in sys.modules:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass

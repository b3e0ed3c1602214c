"""Execution supervisors — the backend seam of the reference, kept call-compatible:

    sup.setup(); sup.cleanup(); sup.config_hash
    sup.call(request, cls_or_fn_name, method_name=None, params=None, distributed_subcall=False)

(kt/serving/execution_supervisor.py:23-157, kt/serving/spmd/spmd_supervisor.py:103-570).
`request.headers` carries X-Request-ID / X-Serialization.  `params` is the call body
{"args": [...], "kwargs": {...}[, "workers": ..., "restart_procs": ...]} — live Python objects on
the local route (nothing is base64/JSON-encoded unless it must cross a process boundary).
"""
from __future__ import annotations

import asyncio
import inspect
import pickle
import threading
from concurrent.futures import FIRST_EXCEPTION, wait
from typing import Any, Dict, List, Optional

from ..distributed import local_pod_ips
from .codec import HTTPException, check_allowed, package_exception, rebuild_exception
from . import fastpickle
from .process_pool import ProcessPool
from .process_worker import instantiate, load_callable, resolve_method, validate_result

DEFAULT_MASTER_PORT = 12345  # kt/serving/spmd/pytorch_process.py:21


class Request:
    """Carrier for the two headers the supervisors read (stands in for starlette's Request)."""

    def __init__(self, headers: Optional[Dict[str, str]] = None):
        self.headers = headers or {}


# ---- per-rank environment contract (process_worker.py:75-102, spmd/*_process.py) ----------------------
def base_env(worker_ips: List[str], node_rank: int, local_rank: int, num_local_procs: int) -> Dict[str, str]:
    return {
        "WORLD_SIZE": str(len(worker_ips) * num_local_procs),
        "RANK": str(node_rank * num_local_procs + local_rank),
        "LOCAL_RANK": str(local_rank),
        "NODE_RANK": str(node_rank),
        "POD_IPS": ",".join(worker_ips),
    }


def framework_env(distribution_type: str, worker_ips, node_rank, local_rank, num_local_procs, port=None):
    env = base_env(worker_ips, node_rank, local_rank, num_local_procs)
    if distribution_type == "pytorch":
        env.update({"MASTER_ADDR": worker_ips[0], "MASTER_PORT": str(port or DEFAULT_MASTER_PORT)})
    elif distribution_type == "jax":  # kt/serving/spmd/jax_process.py:4-44
        env.update({
            "JAX_COORDINATOR_ADDRESS": f"{worker_ips[0]}:{port or 1234}",
            "JAX_NUM_PROCESSES": str(len(worker_ips) * num_local_procs),
            "JAX_PROCESS_ID": str(node_rank * num_local_procs + local_rank),
        })
    return env


def select_worker_nodes(workers_arg, worker_ips: List[str], this_pod_ip: str) -> List[int]:
    """Node indices that take part in the call (semantics of spmd_supervisor.py:219-261)."""
    all_nodes = list(range(len(worker_ips)))
    if not workers_arg:
        return all_nodes
    if isinstance(workers_arg, list):
        targets = set()
        for item in workers_arg:
            if isinstance(item, str) and "." in item:
                if item not in worker_ips:
                    raise ValueError(f"Worker IP '{item}' not found in available workers: {worker_ips}")
                targets.add(worker_ips.index(item))
            elif isinstance(item, int) or (isinstance(item, str) and item.isdigit()):
                idx = int(item)
                if idx < 0 or idx >= len(worker_ips):
                    raise ValueError(f"Worker index {idx} out of range. Valid range: 0-{len(worker_ips)-1}")
                targets.add(idx)
            else:
                raise ValueError(
                    f"Invalid worker specification: {item}. Must be an IP address, integer index, or numeric string."
                )
        return sorted(targets)
    if workers_arg == "any":
        return [worker_ips.index(this_pod_ip)]
    if workers_arg == "ready":
        return all_nodes  # every local rank process is healthy or the pool would have failed
    if isinstance(workers_arg, str):
        me = worker_ips.index(this_pod_ip)
        return sorted({me} | {i for i, ip in enumerate(worker_ips) if workers_arg in ip})
    return all_nodes


def auto_num_processes(distribution_type: str) -> int:
    """num_proc="auto": one rank per visible GPU for pytorch (pytorch_process.py:31-41), else 1."""
    if distribution_type == "pytorch":
        try:
            import torch

            if torch.cuda.is_available():
                return torch.cuda.device_count()
        except ImportError:
            pass
    return 1


def check_callable_name(requested: str, configured: str) -> None:
    """run_callable's guard (kt/serving/http_server.py:1741-1750): the endpoint must name the deployed callable."""
    if configured and requested != configured:
        raise HTTPException(
            404, f"Callable '{requested}' not found in metadata configuration. Found '{configured}' instead")


class ExecutionSupervisor:
    """Non-distributed execution, in-process (BASELINE config C1: "local in-process backend").

    The reference routes to subprocess 0 (execution_supervisor.py:141); here the callable lives in
    the caller's process: sync callables run on the calling thread (so N caller threads give N
    concurrent calls), coroutine functions on one private event loop (so they overlap).
    """

    def __init__(self, pointers=None, init_args=None, name: str = None, allowed_serialization: str = "json,pickle",
                 callable_obj=None, **config):
        self.pointers, self.init_args, self.name = pointers, init_args, name
        self.callable_obj = callable_obj
        self.allowed_serialization = allowed_serialization
        self.config = config
        self.config_hash = hash(repr(sorted(config.items())))
        self._callable = None
        self._loop = None
        self._loop_thread = None

    def setup(self):
        self._callable = instantiate(self.callable_obj, self.init_args) if self.callable_obj is not None \
            else load_callable(self.pointers, self.init_args)

    def cleanup(self):
        self._callable = None
        if self._loop is not None:
            self._loop.call_soon_threadsafe(self._loop.stop)
            self._loop = None

    def _event_loop(self):
        if self._loop is None:
            self._loop = asyncio.new_event_loop()
            self._loop_thread = threading.Thread(target=self._loop.run_forever, name="ktb-local-asyncio", daemon=True)
            self._loop_thread.start()
        return self._loop

    def call(self, request, cls_or_fn_name, method_name=None, params=None, distributed_subcall=False):
        serialization = request.headers.get("X-Serialization", "json")
        if self._callable is None:
            raise HTTPException(503, "Server is loading the callable. Please retry in a moment.")
        check_callable_name(cls_or_fn_name, self.name)
        check_allowed(serialization, self.allowed_serialization)
        params = params or {}
        method = resolve_method(self._callable, cls_or_fn_name, method_name)
        args, kwargs = params.get("args", []), params.get("kwargs", {})
        if inspect.iscoroutinefunction(method):
            result = asyncio.run_coroutine_threadsafe(method(*args, **kwargs), self._event_loop()).result()
        else:
            result = method(*args, **kwargs)
        return validate_result(result, serialization)


class SPMDSupervisor:
    """SPMD fan-out over local rank processes: `workers` emulated pods × `num_proc` ranks each.

    Every rank receives the same (args, kwargs) and the caller gets the rank-ordered list of
    results (spmd_supervisor.py:341,547-570).  The torch-distributed launcher is this class with
    distribution_type="pytorch": MASTER_ADDR resolves to the first local address, MASTER_PORT to
    `port` or 12345, and user code brings up NCCL/gloo itself (e.g. DDP) — the framework only
    launches ranks.
    """

    def __init__(self, distribution_type: str = "spmd", pointers=None, init_args=None, name: str = None,
                 num_proc=None, workers: int = None, quorum_workers: int = None, port: int = None,
                 max_threads_per_proc: int = 10, allowed_serialization: str = "json,pickle",
                 quorum_timeout: int = None, monitor_members: bool = None, restart_procs: bool = True,
                 env_vars: Optional[Dict[str, str]] = None, **extra):
        self.distribution_type = distribution_type or "spmd"
        self.pointers, self.init_args, self.name = pointers, init_args, name
        self.workers = int(workers or quorum_workers or 1)
        if num_proc in (None, 0):
            num_proc = 1
        self.num_proc = auto_num_processes(self.distribution_type) if num_proc == "auto" else int(num_proc)
        self.port = port
        self.max_threads_per_proc = max_threads_per_proc
        self.allowed_serialization = allowed_serialization
        self.env_vars = dict(env_vars or {})
        self.worker_ips = local_pod_ips(self.workers)
        self.pool: Optional[ProcessPool] = None
        cfg = dict(distribution_type=self.distribution_type, workers=self.workers, num_proc=self.num_proc, port=port)
        self.config_hash = hash(repr(sorted(cfg.items())))

    @property
    def world_size(self) -> int:
        return self.workers * self.num_proc

    def setup(self):
        if self.pool is not None:
            self.cleanup()
        pod_names = [f"{self.name}-{r // self.num_proc}" for r in range(self.world_size)]
        self.pool = ProcessPool(
            self.world_size, self.pointers, self.init_args, self.name, max_threads_per_proc=self.max_threads_per_proc,
            base_env=self.env_vars, allowed_serialization=self.allowed_serialization, pod_names=pod_names,
        )

    def cleanup(self):
        if self.pool is not None:
            self.pool.stop()
            self.pool = None

    def rank_envs(self) -> List[Dict[str, str]]:
        envs = []
        for rank in range(self.world_size):
            node, local = divmod(rank, self.num_proc)
            envs.append(framework_env(self.distribution_type, self.worker_ips, node, local, self.num_proc, self.port))
        return envs

    def call(self, request, cls_or_fn_name, method_name=None, params=None, distributed_subcall=False):
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
        payload = fastpickle.dumps((params.get("args", []), params.get("kwargs", {})))  # once for all ranks
        envs = self.rank_envs()
        futures = self.pool.call_all(payload, method_name, envs, serialization, ranks=ranks)
        done, _ = wait(futures, return_when=FIRST_EXCEPTION)
        for f in futures:  # fast-fail in rank order among the finished ones
            if f in done and f.exception() is not None:
                raise f.exception()
        return [pickle.loads(f.result()) for f in futures]

"""CPU restatement of kubetorch's remote-call dispatch path (reference @ 96fac95, v0.5.0).

TEST INFRASTRUCTURE — the parity oracle.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  Nothing under kubetorch_b200/
imports it; the product path fails loudly without the CUDA library.

Parity pinning: this restatement is checked (tests/test_oracle.py) against
  * the reference's own golden vectors for the path (tests/assets/*/inputs.yaml, restated in
    tests/golden/reference_assets.json with file:line provenance), and
  * outputs of the UNMODIFIED reference pod runtime run in the authoring container by
    oracle/make_golden.py (FastAPI TestClient → supervisors → spawned ProcessWorkers),
    committed as tests/golden/ref_runtime.pt: 46 calls on one pod (incl. real gloo DDP / all_reduce ranks) and 10
    calls on TWO real uvicorn pods (127.0.0.1 / 127.0.0.2 x 2 ranks) for the cross-pod half of the path.
Fairness of the timed port: oracle/time_reference.py times the unmodified reference beside OracleRuntime.

Functions and the reference code they follow ("kt/" = python_client/kubetorch/):
  serialize_body          kt/serving/utils.py:730-749      (_serialize_body)
  deserialize_response    kt/serving/utils.py:787-813      (_deserialize_response)
  parse_callable_params   kt/serving/http_server.py:1768-1822
  serialize_result        kt/serving/http_server.py:1825-1842
  status_code_for         kt/serving/http_server.py:1478-1506 (package_exception's status map)
  package_exception       kt/serving/http_server.py:1478-1526
  rehydrate_exception     kt/serving/http_client.py:87-175 (CustomResponse.raise_for_status)
  base_env / pytorch_env  kt/serving/process_worker.py:75-102, kt/serving/spmd/pytorch_process.py:18-29
  select_workers          kt/serving/spmd/spmd_supervisor.py:219-261
  spmd_call               kt/serving/spmd/spmd_supervisor.py:103-570 + process_pool.py:125-234
                          + process_worker.py:109-186 (single pod, P local ranks)
  multipod_call           the coordinator of K pods in the flat topology: spmd_supervisor.py:126-170,219-276,557
                          + remote_worker_pool.py:254-316 (pinned by the two-pod recordings)
  tree_children / fanout_targets   spmd_supervisor.py:68-101 (get_tree_children; pinned by recorded outputs)
  local_call              kt/serving/execution_supervisor.py:105-157 (proc idx 0, bare result)
  OracleRuntime           the same path with real spawned worker processes and queues, used as
                          the timed CPU baseline ("port" kind)
"""
from __future__ import annotations

import asyncio
import base64
import builtins
import importlib
import inspect
import json
import multiprocessing as mp
import os
import pickle
import sys
import threading
import traceback as tb_mod
from contextlib import contextmanager
from typing import Any, Callable, Dict, List, Optional, Tuple

MAGIC_CALL_KWARGS = ["workers", "restart_procs"]  # kt/serving/utils.py:35
DEFAULT_ALLOWED_SERIALIZATION = "json,pickle"
DEFAULT_MASTER_PORT = 12345  # kt/serving/spmd/pytorch_process.py:21


class HTTPException(Exception):
    """Stand-in for fastapi.HTTPException(status_code, detail); str(e) == "<status>: <detail>"."""

    def __init__(self, status_code: int, detail: str):
        super().__init__(f"{status_code}: {detail}")
        self.status_code = status_code
        self.detail = detail


class SerializationError(Exception):
    pass


# ---- client-side codec -----------------------------------------------------------------------------
def build_call_body(*args, **kwargs) -> dict:
    """kt/resources/callables/utils.py:255-261."""
    return {"args": list(args), "kwargs": kwargs}


def serialize_body(body: Optional[dict], serialization: str) -> dict:
    if body is None:
        return {}
    kwargs = body.get("kwargs", {})
    for magic in MAGIC_CALL_KWARGS:  # control kwargs ride outside the serialization boundary
        if magic in kwargs:
            body[magic] = kwargs.pop(magic)
    if serialization == "pickle":
        payload = {"args": body.pop("args"), "kwargs": body.pop("kwargs")}
        body["data"] = base64.b64encode(pickle.dumps(payload)).decode("utf-8")
    return body


def deserialize_response(response_json: Any, serialization: str) -> Any:
    """`response_json` is the already-JSON-decoded HTTP body."""
    if serialization != "pickle":
        return response_json

    def _unwrap(item):
        if isinstance(item, dict) and "data" in item:
            return pickle.loads(base64.b64decode(item["data"].encode("utf-8")))
        return item

    if isinstance(response_json, list):  # SPMD call: list of per-rank envelopes
        return [_unwrap(r) for r in response_json]
    return _unwrap(response_json)


# ---- server-side codec -----------------------------------------------------------------------------
def parse_callable_params(params: Optional[dict], serialization: str, allowed: Optional[str] = None):
    allowed_list = (allowed if allowed is not None else os.getenv("KT_ALLOWED_SERIALIZATION", DEFAULT_ALLOWED_SERIALIZATION)).split(",")
    if serialization not in allowed_list:
        raise HTTPException(400, f"Serialization format '{serialization}' not allowed. Allowed formats: {allowed_list}")
    args, kwargs = [], {}
    if params:
        if serialization == "pickle":
            if isinstance(params, dict) and "data" in params:
                decoded = pickle.loads(base64.b64decode(params.pop("data").encode("utf-8")))
                params.update(decoded)
            elif isinstance(params, str):
                params = pickle.loads(base64.b64decode(params.encode("utf-8")))
        args = params.get("args", [])
        kwargs = params.get("kwargs", {})
    return args, kwargs


def serialize_result(result: Any, serialization: str) -> Any:
    if serialization == "pickle":
        try:
            return {"data": base64.b64encode(pickle.dumps(result)).decode("utf-8")}
        except Exception as e:  # noqa: BLE001
            raise SerializationError(f"Result could not be serialized with pickle: {e}")
    if serialization == "json":
        try:
            json.dumps(result)
        except (TypeError, ValueError) as e:
            raise SerializationError(f"Result could not be serialized to JSON: {e}")
    return result


# ---- error envelope ---------------------------------------------------------------------------------
def status_code_for(exc: BaseException) -> int:
    import concurrent.futures

    if hasattr(exc, "status_code"):
        return exc.status_code
    if isinstance(exc, (TypeError, AssertionError)):
        return 422
    if isinstance(exc, (ValueError, UnicodeError)):  # json.JSONDecodeError is a ValueError
        return 400
    if isinstance(exc, (KeyError, FileNotFoundError)):
        return 404
    if isinstance(exc, PermissionError):
        return 403
    if isinstance(exc, (MemoryError, OSError)):
        return 500
    if isinstance(exc, NotImplementedError):
        return 501
    if isinstance(exc, (asyncio.TimeoutError, concurrent.futures.TimeoutError)):
        return 504
    return 500


def package_exception(exc: BaseException) -> Tuple[int, dict]:
    state = None
    if hasattr(exc, "__getstate__"):
        try:
            state = exc.__getstate__()
            json.dumps(state)
        except Exception:  # noqa: BLE001
            state = None
    envelope = {
        "error_type": exc.__class__.__name__,
        "message": str(exc),
        "traceback": "".join(tb_mod.format_exception(type(exc), exc, exc.__traceback__)),
        "pod_name": os.getenv("POD_NAME", "unknown"),
        "state": state,
    }
    return status_code_for(exc), envelope


def rehydrate_exception(envelope: dict, registry: Optional[dict] = None) -> BaseException:
    """Rebuild the client-side exception the reference raises for a packaged server error."""
    error_type, message = envelope["error_type"], envelope.get("message", "")
    remote_tb, pod_name = envelope["traceback"], envelope["pod_name"]
    state = envelope.get("state") or {}
    registry = registry or {}
    cls = getattr(builtins, error_type, None) or registry.get(error_type)
    exc = None
    if cls is not None:
        try:
            exc = cls.from_dict(state) if (state and hasattr(cls, "from_dict")) else cls(message)
        except Exception:  # noqa: BLE001
            exc = None
    if exc is None:
        exc = type(error_type, (Exception,), {})(message)
    exc.remote_traceback = remote_tb
    exc.pod_name = pod_name

    class RemoteException(exc.__class__):  # str() shows the remote traceback
        def __str__(self):
            return f"{super().__str__()}\n\n{self.remote_traceback.encode().decode('unicode_escape')}"

    wrapped = RemoteException.__new__(RemoteException)
    wrapped.__dict__.update(exc.__dict__)
    wrapped.args = (str(exc),)
    return wrapped


# ---- rank environment --------------------------------------------------------------------------------
def base_env(worker_ips: List[str], node_rank: int, local_rank: int, num_local_procs: int) -> Dict[str, str]:
    return {
        "WORLD_SIZE": str(len(worker_ips) * num_local_procs),
        "RANK": str(node_rank * num_local_procs + local_rank),
        "LOCAL_RANK": str(local_rank),
        "NODE_RANK": str(node_rank),
        "POD_IPS": ",".join(worker_ips),
    }


def pytorch_env(worker_ips, node_rank, local_rank, num_local_procs, port=None) -> Dict[str, str]:
    env = base_env(worker_ips, node_rank, local_rank, num_local_procs)
    env.update({"MASTER_ADDR": worker_ips[0], "MASTER_PORT": str(port or DEFAULT_MASTER_PORT)})
    return env


def rank_env(distribution_type: str, worker_ips, node_rank, local_rank, num_local_procs, port=None):
    if distribution_type == "pytorch":
        return pytorch_env(worker_ips, node_rank, local_rank, num_local_procs, port)
    return base_env(worker_ips, node_rank, local_rank, num_local_procs)


def select_workers(workers_arg, worker_ips: List[str], this_pod_ip: str) -> Tuple[List[str], bool]:
    """Returns (remote_ips_to_call, call_local_procs) for the coordinator."""
    subcall_ips = [ip for ip in worker_ips if ip != this_pod_ip]
    call_local = True
    if not workers_arg:
        return subcall_ips, call_local
    if isinstance(workers_arg, list):
        targets = set()
        for item in workers_arg:
            if isinstance(item, str) and "." in item:
                if item not in worker_ips:
                    raise ValueError(f"Worker IP '{item}' not found in available workers: {worker_ips}")
                targets.add(item)
            elif isinstance(item, int) or (isinstance(item, str) and item.isdigit()):
                idx = int(item)
                if idx < 0 or idx >= len(worker_ips):
                    raise ValueError(f"Worker index {idx} out of range. Valid range: 0-{len(worker_ips)-1}")
                targets.add(worker_ips[idx])
            else:
                raise ValueError(
                    f"Invalid worker specification: {item}. Must be an IP address, integer index, or numeric string."
                )
        subcall_ips = [ip for ip in subcall_ips if ip in targets]
        call_local = this_pod_ip in targets
    elif workers_arg == "any":
        subcall_ips = []
    elif workers_arg == "ready":
        pass
    elif isinstance(workers_arg, str):
        subcall_ips = [ip for ip in subcall_ips if workers_arg in ip]
    return subcall_ips, call_local


def tree_children(sorted_ips: List[str], my_ip: str, fanout: int = 100) -> List[str]:
    """Children of `my_ip` in the self-organising fan-out tree used from 100 pods up
    (kt/serving/spmd/spmd_supervisor.py:68-101): node i's children are indices [i*F+1, i*F+F]."""
    if my_ip not in sorted_ips:
        return []
    first = sorted_ips.index(my_ip) * fanout + 1
    if first >= len(sorted_ips):
        return []
    return sorted_ips[first:min(first + fanout, len(sorted_ips))]


def fanout_targets(worker_ips: List[str], this_ip: str, tree_minimum: int = 100, tree_fanout: int = 50) -> List[str]:
    """Pods the coordinator (or an inner tree node) calls: everyone else when flat, its tree children otherwise
    (kt/serving/spmd/spmd_supervisor.py:178-212)."""
    ips = sorted(worker_ips)
    if len(ips) < tree_minimum:
        return [ip for ip in ips if ip != this_ip]
    return tree_children(ips, this_ip, tree_fanout)


# ---- execution (in-process restatement; sequential over ranks) --------------------------------------
@contextmanager
def _patched_env(env: Dict[str, str]):
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _run_callable(fn: Callable, args, kwargs):
    if inspect.iscoroutinefunction(fn):
        return asyncio.run(fn(*args, **kwargs))
    result = fn(*args, **kwargs)
    if inspect.isawaitable(result):
        result = asyncio.run(_await(result))
    return result


async def _await(x):
    return await x


def _execute_rank(fn, params: dict, serialization: str, env: Dict[str, str], allowed: Optional[str]):
    """One rank's handle_request_async: env → decode → run → encode (or a packaged error)."""
    with _patched_env(env):
        try:
            args, kwargs = parse_callable_params(params, serialization, allowed)
            return serialize_result(_run_callable(fn, args, kwargs), serialization)
        except BaseException as e:  # noqa: BLE001
            return ("__error__",) + package_exception(e)


def _with_status(exc, status: int):
    """Oracle-only: remember the HTTP status the reference server would have answered with."""
    exc.http_status = status
    return exc


def _wire(body: dict) -> dict:
    """The HTTP hop: the body is JSON-encoded by httpx and JSON-decoded by FastAPI."""
    return json.loads(json.dumps(body))


def local_call(fn, *args, serialization: str = "json", allowed: Optional[str] = None, **kwargs):
    """Non-distributed call: routed to subprocess 0 with no distributed env; returns the bare value."""
    body = _wire(serialize_body(build_call_body(*args, **kwargs), serialization))
    res = _execute_rank(fn, body, serialization, {}, allowed)
    if isinstance(res, tuple) and res and res[0] == "__error__":
        raise _with_status(rehydrate_exception(res[2]), res[1])
    return deserialize_response(_wire(res) if serialization != "none" else res, serialization)


def spmd_call(
    fn,
    *args,
    num_proc: int = 1,
    distribution_type: str = "spmd",
    serialization: str = "pickle",
    worker_ips: Optional[List[str]] = None,
    port: Optional[int] = None,
    allowed: Optional[str] = None,
    **kwargs,
) -> List[Any]:
    """Single pod × num_proc ranks. The same params go to every rank; result = rank-ordered list."""
    worker_ips = worker_ips or ["localhost"]
    body = _wire(serialize_body(build_call_body(*args, **kwargs), serialization))
    workers_arg = body.get("workers")
    try:
        select_workers(workers_arg, worker_ips, worker_ips[0])  # raises the reference's selector errors
    except ValueError as e:
        status, envelope = package_exception(e)
        raise _with_status(rehydrate_exception(envelope), status)
    responses = []
    for local_rank in range(num_proc):
        # mp.Queue pickles the params dict once per rank: each rank decodes its own copy
        params = pickle.loads(pickle.dumps(body))
        env = rank_env(distribution_type, worker_ips, 0, local_rank, num_proc, port)
        res = _execute_rank(fn, params, serialization, env, allowed)
        if isinstance(res, tuple) and res and res[0] == "__error__":
            raise _with_status(rehydrate_exception(res[2]), res[1])  # fast-fail on the first failing rank
        responses.append(res)
    return deserialize_response(_wire(responses), serialization)


def multipod_call(
    fn,
    *args,
    num_proc: int = 1,
    pod_ips: List[str],
    distribution_type: str = "spmd",
    serialization: str = "pickle",
    port: Optional[int] = None,
    allowed: Optional[str] = None,
    **kwargs,
) -> List[Any]:
    """K pods x num_proc ranks in the flat topology (fewer pods than tree_minimum), called on pod_ips[0].

    kt/serving/spmd/spmd_supervisor.py:126-170 (coordinator: sort the discovered IPs, move itself to the front, publish
    them as POD_IPS), :219-261 (`workers=` narrows the remote pods and may exclude the coordinator's own ranks),
    :270-276 (node_rank = position in that list), :341-365 (same params to every local rank), remote pods run the same
    supervisor as a distributed_subcall with the coordinator's POD_IPS (remote_worker_pool.py:254-316), and
    :557 `responses = local_responses + worker_responses`: the coordinator's ranks first, then the pods in list order."""
    this_pod_ip = pod_ips[0]
    worker_ips = sorted(pod_ips)
    worker_ips.remove(this_pod_ip)
    worker_ips.insert(0, this_pod_ip)
    body = _wire(serialize_body(build_call_body(*args, **kwargs), serialization))
    try:
        subcall_ips, call_local = select_workers(body.get("workers"), worker_ips, this_pod_ip)
    except ValueError as e:
        status, envelope = package_exception(e)
        raise _with_status(rehydrate_exception(envelope), status)
    responses = []
    for ip in ([this_pod_ip] if call_local else []) + list(subcall_ips):
        node_rank = worker_ips.index(ip)
        pod_params = body if ip == this_pod_ip else _wire(body)     # one more JSON hop to a remote pod
        pod_responses = []
        for local_rank in range(num_proc):
            params = pickle.loads(pickle.dumps(pod_params))           # mp.Queue pickles the params once per rank
            env = rank_env(distribution_type, worker_ips, node_rank, local_rank, num_proc, port)
            res = _execute_rank(fn, params, serialization, env, allowed)
            if isinstance(res, tuple) and res and res[0] == "__error__":
                raise _with_status(rehydrate_exception(res[2]), res[1])
            pod_responses.append(res)
        responses.extend(pod_responses if ip == this_pod_ip else _wire(pod_responses))
    return deserialize_response(_wire(responses), serialization)


# ---- the same path with real processes (timed CPU baseline) -----------------------------------------
def _oracle_worker(idx: int, req_q, resp_q, module_name: str, fn_name: str, extra_path: str):
    if extra_path and extra_path not in sys.path:
        sys.path.insert(0, extra_path)
    fn = getattr(importlib.import_module(module_name), fn_name)
    while True:
        req = req_q.get()
        if req == "SHUTDOWN":
            break
        os.environ.update(req["env"])
        try:
            args, kwargs = parse_callable_params(req["params"], req["serialization"], None)
            out = serialize_result(_run_callable(fn, args, kwargs), req["serialization"])
        except BaseException as e:  # noqa: BLE001
            out = {"__error__": package_exception(e)}
        resp_q.put({"id": req["id"], "idx": idx, "result": out})


class OracleRuntime:
    """Spawned worker per rank, one request queue each, one shared response queue — the reference's
    ProcessPool shape (kt/serving/process_pool.py:14-69) — driven by the client codec and a JSON
    round-trip standing in for the HTTP hop. Used only to time the CPU path."""

    def __init__(self, module_name: str, fn_name: str, num_proc: int, distribution_type: str = "spmd",
                 extra_path: Optional[str] = None):
        self.num_proc = num_proc
        self.distribution_type = distribution_type
        ctx = mp.get_context("spawn")
        self.resp_q = ctx.Queue()
        self.req_qs = [ctx.Queue() for _ in range(num_proc)]
        extra_path = extra_path or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        self.procs = [
            ctx.Process(target=_oracle_worker, args=(i, self.req_qs[i], self.resp_q, module_name, fn_name, extra_path),
                        daemon=True)
            for i in range(num_proc)
        ]
        for p in self.procs:
            p.start()
        self._next = 0

    def call(self, *args, serialization: str = "pickle", **kwargs) -> List[Any]:
        body = serialize_body(build_call_body(*args, **kwargs), serialization)  # client: pack
        wire = json.dumps(body)  # httpx encodes
        params = json.loads(wire)  # FastAPI decodes
        rid = self._next
        self._next += 1
        for i, q in enumerate(self.req_qs):  # coordinator → every local rank (pickled per rank by mp.Queue)
            env = rank_env(self.distribution_type, ["localhost"], 0, i, self.num_proc)
            q.put({"id": rid, "params": params, "serialization": serialization, "env": env})
        got: Dict[int, Any] = {}
        while len(got) < self.num_proc:
            r = self.resp_q.get()
            if r["id"] == rid:
                got[r["idx"]] = r["result"]
        ordered = [got[i] for i in range(self.num_proc)]
        for r in ordered:
            if isinstance(r, dict) and "__error__" in r:
                raise rehydrate_exception(r["__error__"][1])
        response = json.loads(json.dumps(ordered))  # FastAPI encodes the list, client decodes
        return deserialize_response(response, serialization)

    def close(self):
        for q in self.req_qs:
            try:
                q.put("SHUTDOWN")
            except Exception:  # noqa: BLE001
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

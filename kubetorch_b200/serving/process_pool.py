"""Pool of rank worker processes — the local scatter/gather of a call.

Role of kt/serving/process_pool.py:14-255 (ProcessPool), re-designed: one duplex pipe per worker,
one reader thread multiplexing all pipes with multiprocessing.connection.wait, futures instead of
Event+dict, and the call payload pickled ONCE for all ranks.  Worker death fails the pending
futures with PodTerminatedError (the reference's analogue of a dead pod, serving/utils.py:111-190).
"""
from __future__ import annotations

import itertools
import multiprocessing as mp
import pickle
import threading
from concurrent.futures import Future
from multiprocessing.connection import wait as mp_wait
from typing import Dict, List, Optional

from ..exceptions import PodTerminatedError, StartupError
from .codec import rebuild_exception
from .process_worker import SHUTDOWN, worker_main


class ProcessPool:
    def __init__(self, num_processes: int, pointers, init_args, name: str, max_threads_per_proc: int = 10,
                 base_env: Optional[Dict[str, str]] = None, allowed_serialization: str = "json,pickle",
                 pod_names: Optional[List[str]] = None, start_timeout: float = 300.0,
                 gpu_cfgs: Optional[List[dict]] = None):
        self.num_processes = int(num_processes)
        self.name = name
        self._ctx = mp.get_context("spawn")  # like the reference (execution_supervisor.py:66-67)
        self._conns = []
        self._send_locks: List[threading.Lock] = []   # one writer at a time per pipe (header+payload are two writes)
        self._procs = []
        self._pending: Dict[int, Future] = {}
        self._pending_owner: Dict[int, int] = {}
        self._lock = threading.Lock()
        self._ids = itertools.count()
        self._closed = False
        self.full_replies = bool(gpu_cfgs)  # GPU ranks attach arena metadata to their replies
        self.pod_names = pod_names or [f"{name}-rank{i}" for i in range(self.num_processes)]
        ready: List[Future] = []
        for i in range(self.num_processes):
            parent, child = self._ctx.Pipe(duplex=True)
            env = dict(base_env or {})
            env.setdefault("POD_NAME", self.pod_names[i])
            p = self._ctx.Process(
                target=worker_main,
                args=(child, i, pointers, init_args, name, max_threads_per_proc, env, allowed_serialization,
                      gpu_cfgs[i] if gpu_cfgs else None),
                daemon=True, name=f"ktb-worker-{name}-{i}",
            )
            p.start()
            child.close()
            self._conns.append(parent)
            self._send_locks.append(threading.Lock())
            self._procs.append(p)
        self._reader = threading.Thread(target=self._read_loop, name=f"ktb-pool-{name}", daemon=True)
        # startup handshake: each worker reports the outcome of importing the callable (id -1)
        self._startup: Dict[int, Future] = {i: Future() for i in range(self.num_processes)}
        self._reader.start()
        for i in range(self.num_processes):
            try:
                self._startup[i].result(timeout=start_timeout)
            except BaseException as e:
                self.stop()
                if isinstance(e, TimeoutError):
                    raise StartupError(f"worker {i} of '{name}' did not start within {start_timeout}s")
                raise

    def __len__(self):
        return self.num_processes

    # ---- response routing -------------------------------------------------------------------------
    def _read_loop(self):
        live = {c: i for i, c in enumerate(self._conns)}
        while live and not self._closed:
            for conn in mp_wait(list(live), timeout=0.5):
                idx = live[conn]
                try:
                    msg = pickle.loads(conn.recv_bytes())
                except (EOFError, OSError):
                    del live[conn]
                    self._fail_worker(idx)
                    continue
                if msg["id"] == -1:
                    fut = self._startup[idx]
                    if msg["ok"]:
                        fut.set_result(True)
                    else:
                        fut.set_exception(rebuild_exception(msg["envelope"]))
                    continue
                with self._lock:
                    fut = self._pending.pop(msg["id"], None)
                    self._pending_owner.pop(msg["id"], None)
                if fut is None:
                    continue
                if msg["ok"]:
                    fut.set_result(msg if self.full_replies else msg["result"])
                else:
                    fut.set_exception(rebuild_exception(msg["envelope"]))

    def _fail_worker(self, idx: int):
        if self._closed:
            return
        err = PodTerminatedError(pod_name=self.pod_names[idx], reason="WorkerProcessExited", status_code=503)
        if not self._startup[idx].done():
            self._startup[idx].set_exception(err)
        with self._lock:
            dead = [rid for rid, owner in self._pending_owner.items() if owner == idx]
            futs = [self._pending.pop(rid) for rid in dead]
            for rid in dead:
                self._pending_owner.pop(rid, None)
        for f in futs:
            f.set_exception(err)

    # ---- calls ----------------------------------------------------------------------------------------
    def submit(self, idx: int, payload: bytes, method_name: Optional[str], env: Dict[str, str],
               serialization: str, extra: Optional[dict] = None) -> Future:
        """Send one request to worker idx. `payload` = pickle.dumps((args, kwargs)) made once by the caller."""
        fut: Future = Future()
        rid = next(self._ids)
        if not self._procs[idx].is_alive():
            fut.set_exception(PodTerminatedError(pod_name=self.pod_names[idx], reason="WorkerProcessExited"))
            return fut
        with self._lock:
            self._pending[rid] = fut
            self._pending_owner[rid] = idx
        req = {"id": rid, "payload": payload, "method": method_name, "env": env, "serialization": serialization}
        if extra:
            req.update(extra)
        data = pickle.dumps(req, protocol=5)
        try:
            with self._send_locks[idx]:   # concurrent callers (threads, async) must not interleave frames on one pipe
                self._conns[idx].send_bytes(data)
        except (OSError, ValueError):
            self._fail_worker(idx)
        return fut

    def call_all(self, payload: bytes, method_name: Optional[str], envs: List[Dict[str, str]], serialization: str,
                 ranks: Optional[List[int]] = None, extras: Optional[Dict[int, dict]] = None) -> List[Future]:
        ranks = list(range(self.num_processes)) if ranks is None else ranks
        return [self.submit(i, payload, method_name, envs[i], serialization, (extras or {}).get(i)) for i in ranks]

    def stop(self):
        if self._closed:
            return
        self._closed = True
        for c, lk in zip(self._conns, self._send_locks):
            try:
                with lk:
                    c.send_bytes(SHUTDOWN)
            except Exception:  # noqa: BLE001
                pass
        for p in self._procs:
            p.join(timeout=3)
            if p.is_alive():
                p.terminate()
                p.join(timeout=2)
            if p.is_alive():
                p.kill()
        for c in self._conns:
            try:
                c.close()
            except Exception:  # noqa: BLE001
                pass
        err = PodTerminatedError(pod_name=self.name, reason="PoolStopped")
        with self._lock:
            futs = list(self._pending.values())
            self._pending.clear()
            self._pending_owner.clear()
        for f in futs:
            if not f.done():
                f.set_exception(err)

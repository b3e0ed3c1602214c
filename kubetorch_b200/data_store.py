"""kt.put / kt.get / kt.ls / kt.rm for GPU tensors and state dicts on one box.

Reference path: kt/data_store/data_store_cmds.py:23-296 → GPUTransferManager.publish/retrieve
(kt/data_store/gpu_transfer.py:164-559) → per-node PodDataServer holding CUDA-IPC handles and running
NCCL broadcasts (kt/data_store/pod_data_server.py:405-579,1172-1341); packed mode = torch.cat → one
NCCL broadcast → copy_ loop (gpu_transfer.py:291-359,482-559).

Here the closed metadata server is replaced by an in-process registry (one controller process drives the
local GPUs) and the data path by libktb200 kernels:

  put              registers the source tensors (zero-copy, like locale="local") + a CUDA event
  get              one segmented kernel on the destination GPU pulls every leaf from the source GPU over
                   NVLink (ktb_map_batch identity with peer sources) — no staging
  BroadcastWindow  putters and getters join a quorum (world_size / timeout); the transfer then runs once:
                   pack=True  → ktb_pack (leaves → arena) → ktb_broadcast (one read, N-1 peer stores into
                                every getter GPU's arena) → ktb_unpack on each getter GPU
                   pack=False → one segmented pull per getter
Filesystem keys (rsync store) are a Kubernetes feature and raise NotImplementedError.

Across rank PROCESSES (the reference's real usage: one pod puts, another gets): when KTB_STORE_DIR is set — the GPU
SPMD supervisor sets it for its rank processes — `put` additionally publishes each leaf through a library arena
(one device copy + CUDA IPC handle + a small descriptor file under KTB_STORE_DIR, the stand-in for the metadata
server), and `get` in another process opens the handle and pulls with the same segmented kernel.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple, Union

import hashlib
import json
import os

from .exceptions import DataStoreError


@dataclass
class BroadcastWindow:
    """Same fields and validation as kt/data_store/types.py:23-107."""

    timeout: Optional[float] = None
    world_size: Optional[int] = None
    ips: Optional[List[str]] = None
    group_id: Optional[str] = None
    fanout: Optional[int] = None
    pack: bool = False

    def __post_init__(self):
        if self.timeout is None and self.world_size is None and self.ips is None:
            raise ValueError("BroadcastWindow requires at least one of: timeout, world_size, or ips")

    def to_dict(self) -> dict:
        return {"timeout": self.timeout, "world_size": self.world_size, "ips": self.ips, "group_id": self.group_id,
                "fanout": self.fanout, "pack": self.pack}


# ---- helpers mirrored from gpu_transfer.py:73-110 ----------------------------------------------------------
def _is_gpu_tensor(obj) -> bool:
    try:
        import torch
    except ImportError:
        return False
    return isinstance(obj, torch.Tensor) and obj.is_cuda


def _is_gpu_data(obj) -> bool:
    if _is_gpu_tensor(obj):
        return True
    if isinstance(obj, dict):
        return any(_is_gpu_tensor(v) or (isinstance(v, dict) and _is_gpu_data(v)) for v in obj.values())
    return False


def _flatten_state_dict(state_dict: Dict, prefix: str = "") -> Dict[str, Any]:
    out = {}
    for k, v in state_dict.items():
        full = f"{prefix}.{k}" if prefix else k
        if isinstance(v, dict):
            out.update(_flatten_state_dict(v, full))
        else:
            out[full] = v
    return out


def _leaves(data, what: str) -> List[Tuple[str, Any]]:
    """Sorted (tensor_key, tensor) pairs; '' for a bare tensor. Validation messages follow the reference."""
    import torch

    if isinstance(data, torch.Tensor):
        if not data.is_cuda:
            raise ValueError("Tensor must be on a CUDA device" if what == "src"
                             else "Destination tensor must be on a CUDA device")
        return [("", data)]
    if isinstance(data, dict):
        out = []
        for k, v in sorted(_flatten_state_dict(data).items()):
            if isinstance(v, torch.Tensor):
                if not v.is_cuda:
                    raise ValueError(f"Tensor at '{k}' must be on a CUDA device")
                out.append((k, v))
        return out
    raise ValueError("Data must be a torch.Tensor or dict of tensors" if what == "src"
                     else "dest must be a torch.Tensor or dict of tensors")


class _Entry:
    __slots__ = ("tensor", "event", "stamp")

    def __init__(self, tensor, event):
        self.tensor, self.event, self.stamp = tensor, event, time.time()


class _Group:
    def __init__(self):
        self.cond = threading.Condition()
        self.putters: List[Tuple[str, List[Tuple[str, Any]]]] = []
        self.getters: List[Tuple[str, List[Tuple[str, Any]]]] = []
        self.done = False
        self.error: Optional[BaseException] = None
        self.closed = False


_lock = threading.Lock()
_registry: Dict[str, _Entry] = {}
_groups: Dict[str, _Group] = {}


# ---- cross-process registry (descriptor files + CUDA IPC arenas) ---------------------------------------------
_shared_arenas: Dict[str, Any] = {}     # full_key -> Arena owned by this process
_opened: Dict[str, int] = {}            # handle hex -> mapped device pointer in this process
_opened_by_key: Dict[str, str] = {}     # full_key -> handle hex this process has mapped for it
_retired_arenas: List[Any] = []         # outgrown arenas: other processes may still map them, so they are not freed


def _store_dir() -> Optional[str]:
    d = os.environ.get("KTB_STORE_DIR")
    if d:
        os.makedirs(d, exist_ok=True)
    return d or None


def _desc_path(d: str, full_key: str) -> str:
    return os.path.join(d, hashlib.sha1(full_key.encode()).hexdigest() + ".json")


def _publish_shared(full_key: str, t) -> None:
    """Copy the leaf into an IPC-exportable arena and write its descriptor (visible to other rank processes)."""
    import torch

    from .device import ops

    d = _store_dir()
    if d is None:
        return
    dev = t.device.index
    nbytes = max(t.numel() * t.element_size(), 1)
    old = _shared_arenas.pop(full_key, None)
    if old is not None and old.device == dev and old.nbytes >= nbytes:
        arena = old              # re-put of the same key: overwrite in place, the exported handle stays valid
    else:
        if old is not None:
            # freeing exported memory that another process still maps is undefined behaviour (CUDA IPC) and there is
            # no acknowledgement channel between getters and putters: keep it until kt.rm / process exit
            _retired_arenas.append(old)
        arena = ops.Arena(dev, nbytes)
    with torch.cuda.device(dev):
        if t.numel():
            ops.map_tensor(t.contiguous().reshape(-1).view(torch.uint8), "identity",
                           out=arena.tensor(torch.uint8, t.numel() * t.element_size()))
        torch.cuda.current_stream(dev).synchronize()     # bytes are in HBM before the descriptor becomes visible
    _shared_arenas[full_key] = arena
    desc = {"key": full_key, "dtype": str(t.dtype).replace("torch.", ""), "shape": list(t.shape), "device": dev,
            "nbytes": t.numel() * t.element_size(), "handle": arena.export().hex(), "pid": os.getpid()}
    tmp = _desc_path(d, full_key) + f".{os.getpid()}.tmp"
    with open(tmp, "w") as f:
        json.dump(desc, f)
    os.replace(tmp, _desc_path(d, full_key))


def _lookup_shared(full_key: str):
    """A (tensor view, None) of another process's published leaf, or None."""
    import torch

    from .device import ops

    d = _store_dir()
    if d is None or not os.path.exists(_desc_path(d, full_key)):
        return None
    with open(_desc_path(d, full_key)) as f:
        desc = json.load(f)
    if desc["pid"] == os.getpid():
        return None
    dev = torch.cuda.current_device()
    ptr = _opened.get(desc["handle"])
    if ptr is None:
        stale = _opened_by_key.get(full_key)
        if stale is not None and stale in _opened:       # the putter replaced the arena: drop our mapping of the old one
            torch.cuda.synchronize(dev)
            try:
                ops.ipc_close(dev, _opened.pop(stale))
            except Exception:  # noqa: BLE001
                pass
        ptr = _opened[desc["handle"]] = ops.ipc_open(dev, bytes.fromhex(desc["handle"]))
        _opened_by_key[full_key] = desc["handle"]
    holder = type("_CAI", (), {})()
    holder.__cuda_array_interface__ = {"shape": (max(desc["nbytes"], 1),), "typestr": "|u1", "data": (ptr, False),
                                       "version": 3}
    flat = torch.as_tensor(holder, device=f"cuda:{dev}")[:desc["nbytes"]]
    dtype = getattr(torch, desc["dtype"])
    return flat.view(dtype).reshape(desc["shape"]) if desc["nbytes"] else torch.empty(desc["shape"], dtype=dtype,
                                                                                        device=f"cuda:{dev}")


def _full_key(key: str, tensor_key: str) -> str:
    return f"{key}/{tensor_key}" if tensor_key else key


def _check_pair(full_key: str, src, dst):
    if src.dtype != dst.dtype or src.numel() != dst.numel():
        raise ValueError(
            f"Destination for '{full_key}' is {tuple(dst.shape)} {dst.dtype} but the stored tensor is "
            f"{tuple(src.shape)} {src.dtype}")
    if not dst.is_contiguous():
        raise ValueError(f"Destination for '{full_key}' must be contiguous")


def _pull(pairs: List[Tuple[Any, Any, Any]]):
    """pairs: (src tensor, src event, dst tensor). One segmented launch per (dst device, dtype size class)."""
    import torch

    from .device import ops

    by_dev: Dict[int, List] = {}
    for src, ev, dst in pairs:
        by_dev.setdefault(dst.device.index, []).append((src.contiguous(), ev, dst))
    for dev, items in by_dev.items():
        ops.ensure_init({dev} | {s.device.index for s, _, _ in items})
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev)
            for _, ev, _ in items:
                if ev is not None:
                    st.wait_event(ev)          # the putter's producing work is done
            srcs = [s.reshape(-1).view(torch.uint8) for s, _, _ in items]
            dsts = [d.reshape(-1).view(torch.uint8) for _, _, d in items]
            ops.map_batch(srcs, "identity", outs=dsts, device=dev)   # runs on the getter's GPU, peer sources


def _packed_broadcast(src_leaves: List[Any], events: List[Any], getter_leaves: List[List[Any]]):
    """pack → broadcast → unpack: one read of the source, one peer store per getter GPU."""
    import torch

    from .device import ops

    src_dev = src_leaves[0].device.index
    dst_devs = [g[0].device.index for g in getter_leaves]
    ops.ensure_init({src_dev} | set(dst_devs))
    with torch.cuda.device(src_dev):
        st = torch.cuda.current_stream(src_dev)
        for ev in events:
            if ev is not None:
                st.wait_event(ev)
        arena, offsets = ops.pack([t.contiguous() for t in src_leaves])
        total = arena.numel()
        remote = {}
        for dev in set(dst_devs):
            if dev != src_dev:
                remote[dev] = torch.empty(total, dtype=torch.uint8, device=f"cuda:{dev}")
        if remote:
            ops.broadcast(arena, list(remote.values()))
        ready = torch.cuda.Event()
        ready.record(st)
    for dev, leaves in zip(dst_devs, getter_leaves):
        with torch.cuda.device(dev):
            torch.cuda.current_stream(dev).wait_event(ready)
            ops.unpack(arena if dev == src_dev else remote[dev], offsets, leaves)
    # keep the staging arenas alive until the consumers' streams have used them
    for dev in set(dst_devs):
        torch.cuda.current_stream(dev).synchronize()


# ---- public API -----------------------------------------------------------------------------------------------
def put(key: Union[str, List[str]], src=None, locale: str = "store", broadcast: Optional[BroadcastWindow] = None,
        contents: bool = False, filter_options: Optional[str] = None, force: bool = False, verbose: bool = False,
        namespace: Optional[str] = None, kubeconfig_path: Optional[str] = None, start_rsyncd: bool = True,
        base_path: str = "/", nccl_port: int = 29500, nccl_pg_mode: Optional[str] = None):
    if src is None:
        raise ValueError("src is required. Provide a path for filesystem data or a GPU tensor/dict for GPU data.")
    if not _is_gpu_data(src):
        raise NotImplementedError("filesystem keys use the Kubernetes rsync store; this backend handles GPU data")
    if isinstance(key, list):
        raise ValueError("GPU data transfer only supports a single key, not a list of keys.")
    import torch

    leaves = _leaves(src, "src")
    with _lock:
        for tk, t in leaves:
            ev = torch.cuda.Event()
            with torch.cuda.device(t.device):
                ev.record(torch.cuda.current_stream(t.device))
            _registry[_full_key(key, tk)] = _Entry(t, ev)
    if _store_dir() is not None:
        for tk, t in leaves:
            _publish_shared(_full_key(key, tk), t)
    if broadcast is not None:
        return _join(key, leaves, broadcast, role="put")
    return None


def get(key: Union[str, List[str]], dest=None, broadcast: Optional[BroadcastWindow] = None, contents: bool = False,
        filter_options: Optional[str] = None, force: bool = False, verbose: bool = False,
        namespace: Optional[str] = None, kubeconfig_path: Optional[str] = None, nccl_pg_mode: Optional[str] = None):
    if dest is None or not _is_gpu_data(dest):
        raise NotImplementedError("filesystem keys use the Kubernetes rsync store; pass a GPU tensor or dict as dest")
    if isinstance(key, list):
        raise ValueError("GPU data transfer only supports a single key, not a list of keys.")
    leaves = _leaves(dest, "dest")
    if broadcast is not None:
        return _join(key, leaves, broadcast, role="get")
    pairs = []
    with _lock:
        for tk, d in leaves:
            fk = _full_key(key, tk)
            ent = _registry.get(fk)
            if ent is None:
                remote = _lookup_shared(fk)           # published by another rank process
                if remote is None:
                    raise DataStoreError(f"Key '{fk}' not found in the GPU data store")
                _check_pair(fk, remote, d)
                pairs.append((remote, None, d))
                continue
            _check_pair(fk, ent.tensor, d)
            pairs.append((ent.tensor, ent.event, d))
    _pull(pairs)
    return None


def ls(key: str = "", verbose: bool = False, namespace: Optional[str] = None, **_) -> List[str]:
    with _lock:
        keys = {k for k in _registry if k.startswith(key)}
    d = _store_dir()
    if d is not None:
        for name in os.listdir(d):
            if name.endswith(".json"):
                try:
                    with open(os.path.join(d, name)) as f:
                        k = json.load(f)["key"]
                    if k.startswith(key):
                        keys.add(k)
                except (OSError, ValueError, KeyError):
                    pass
    return sorted(keys)


def rm(key: str, recursive: bool = False, verbose: bool = False, namespace: Optional[str] = None, **_) -> None:
    with _lock:
        victims = [k for k in _registry if k == key or k.startswith(key + "/")]
        shared = [k for k in list(_shared_arenas) if k == key or k.startswith(key + "/")]
        if not victims and not shared:
            raise DataStoreError(f"Key '{key}' not found in the GPU data store")
        for k in victims:
            del _registry[k]
        d = _store_dir()
        for k in shared:
            _shared_arenas.pop(k).free()
            if d is not None and os.path.exists(_desc_path(d, k)):
                os.remove(_desc_path(d, k))


# ---- BroadcastWindow quorum (threads of the controller process) ------------------------------------------------
def _join_shared(key: str, leaves, bw: BroadcastWindow, role: str):
    """BroadcastWindow ACROSS rank processes (the reference's real usage: every pod joins the same window,
    kt/data_store/pod_data_server.py:1172-1341).  The rendezvous is a directory under KTB_STORE_DIR (the stand-in for
    the metadata server): every participant drops a join file, the window closes when `world_size` joined (or, for a
    timeout-only window, at the deadline with at least one putter and one getter); getters then pull the putter's
    leaves out of its IPC-exported arenas with the segmented kernel on their own GPU and drop a done file, the putter
    returns once every getter is done.  A window that cannot close raises DataStoreError on every participant and
    leaves no state behind, so the next window with the same group id works (the reference's fault-injection case,
    tests/assets/kv_store/gpu_helper.py:607-670)."""
    import uuid

    import torch

    gid = bw.group_id or f"auto:{key}"
    gdir = os.path.join(_store_dir(), "bw_" + hashlib.sha1(gid.encode()).hexdigest())
    os.makedirs(gdir, exist_ok=True)
    me = f"{time.time_ns():020d}_{role}_{os.getpid()}_{uuid.uuid4().hex[:8]}"
    mine = os.path.join(gdir, me + ".join")
    with open(mine + ".tmp", "w") as f:
        json.dump({"role": role, "key": key, "leaves": [tk for tk, _ in leaves], "pid": os.getpid()}, f)
    os.replace(mine + ".tmp", mine)
    deadline = time.time() + (bw.timeout if bw.timeout is not None else 600.0)

    def members():
        out = []
        for name in sorted(os.listdir(gdir)):
            if name.endswith(".join"):
                try:
                    with open(os.path.join(gdir, name)) as f:
                        out.append((name[:-5], json.load(f)))
                except (OSError, ValueError):
                    pass
        return out

    def leave():
        for suffix in (".join", ".done"):
            try:
                os.remove(os.path.join(gdir, me + suffix))
            except OSError:
                pass

    closed = os.path.join(gdir, "CLOSED")
    while True:
        mem = members()
        n_put = sum(1 for _, m in mem if m["role"] == "put")
        n_get = len(mem) - n_put
        if os.path.exists(closed):
            break
        if bw.world_size is not None and len(mem) >= bw.world_size:
            break
        if time.time() >= deadline:
            if bw.world_size is None and n_put and n_get:
                break                                    # a timeout-only window closes at its deadline
            leave()
            raise DataStoreError(f"BroadcastWindow '{gid}' timed out with {n_put} putter(s) and {n_get} getter(s)")
        time.sleep(0.002)
    try:
        open(closed, "a").close()                        # late joiners of THIS window see it closed and proceed
        mem = members()
        names = [name for name, _ in mem]
        rank = names.index(me) if me in names else len(names)
        if not any(m["role"] == "put" for _, m in mem):
            raise DataStoreError("BroadcastWindow closed without a putter")
        getters = [name for name, m in mem if m["role"] == "get"]
        if role == "get":
            pairs = []
            for tk, dst in leaves:
                fk = _full_key(key, tk)
                with _lock:
                    ent = _registry.get(fk)
                src = (ent.tensor, ent.event) if ent is not None else (_lookup_shared(fk), None)
                if src[0] is None:
                    raise DataStoreError(f"Key '{fk}' was not published in this broadcast group")
                _check_pair(fk, src[0], dst)
                pairs.append((src[0], src[1], dst))
            _pull(pairs)
            for dev in {d.device.index for _, d in leaves}:
                torch.cuda.synchronize(dev)              # the bytes are in the destination before "done" is visible
            open(os.path.join(gdir, me + ".done"), "a").close()
        else:
            while True:                                  # the source must stay valid until every getter has pulled
                done = {n[:-5] for n in os.listdir(gdir) if n.endswith(".done")}
                if all(g in done for g in getters):
                    break
                if time.time() >= deadline + 30.0:
                    raise DataStoreError(f"BroadcastWindow '{gid}': getters did not finish pulling")
                time.sleep(0.002)
        return {"rank": rank, "world_size": len(mem), "group_id": gid, "role": role}
    finally:
        if role == "put":                                # the putter is the last one out: clear the window's state
            for name in os.listdir(gdir):
                try:
                    os.remove(os.path.join(gdir, name))
                except OSError:
                    pass
        elif not os.path.exists(os.path.join(gdir, me + ".done")):
            leave()


def _join(key: str, leaves, bw: BroadcastWindow, role: str):
    if _store_dir() is not None:
        return _join_shared(key, leaves, bw, role)
    gid = bw.group_id or f"auto:{key}"
    with _lock:
        grp = _groups.get(gid)
        if grp is None or grp.closed:
            grp = _groups[gid] = _Group()
    deadline = time.time() + (bw.timeout if bw.timeout is not None else 600.0)
    with grp.cond:
        (grp.putters if role == "put" else grp.getters).append((key, leaves))
        rank = len(grp.putters) + len(grp.getters) - 1
        n = len(grp.putters) + len(grp.getters)
        quorum = bw.world_size is not None and n >= bw.world_size
        if quorum:
            _run_group(grp, bw)
            grp.cond.notify_all()
        else:
            while not grp.done and grp.error is None:
                left = deadline - time.time()
                if left <= 0:
                    if bw.world_size is None and grp.putters and grp.getters:   # timeout-only window closes here
                        _run_group(grp, bw)
                        grp.cond.notify_all()
                        break
                    grp.error = DataStoreError(
                        f"BroadcastWindow '{gid}' timed out with {len(grp.putters)} putter(s) and "
                        f"{len(grp.getters)} getter(s)")
                    grp.closed = True
                    grp.cond.notify_all()
                    break
                grp.cond.wait(timeout=min(left, 0.25))
        if grp.error is not None:
            raise grp.error
        world = len(grp.putters) + len(grp.getters)
    return {"rank": rank, "world_size": world, "group_id": gid, "role": role}


def _run_group(grp: _Group, bw: BroadcastWindow):
    """Called with grp.cond held by the participant that completed the quorum."""
    try:
        if not grp.putters:
            raise DataStoreError("BroadcastWindow closed without a putter")
        src: Dict[str, Any] = {}
        for key, leaves in grp.putters:
            for tk, t in leaves:
                src[_full_key(key, tk)] = t
        with _lock:
            events = {fk: (_registry[fk].event if fk in _registry else None) for fk in src}
        getter_sets = []
        for key, leaves in grp.getters:
            pairs = []
            for tk, d in leaves:
                fk = _full_key(key, tk)
                if fk not in src:
                    raise DataStoreError(f"Key '{fk}' was not published in this broadcast group")
                _check_pair(fk, src[fk], d)
                pairs.append((fk, d))
            getter_sets.append(pairs)
        same_keys = len({tuple(fk for fk, _ in p) for p in getter_sets}) == 1
        src_devs = {t.device.index for t in src.values()}
        if bw.pack and getter_sets and same_keys and len(src_devs) == 1 and len(getter_sets[0]) > 1:
            order = [fk for fk, _ in getter_sets[0]]
            _packed_broadcast([src[fk] for fk in order], [events[fk] for fk in order],
                              [[d for _, d in p] for p in getter_sets])
        else:
            _pull([(src[fk], events[fk], d) for p in getter_sets for fk, d in p])
        grp.done = True
    except BaseException as e:  # noqa: BLE001
        grp.error = e
    finally:
        grp.closed = True

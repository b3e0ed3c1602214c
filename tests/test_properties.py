"""not-gpu: property tests (hypothesis) of the host arithmetic and codecs against the oracle restatement."""
import ctypes
import json

import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from kubetorch_b200.device import lib as L
from kubetorch_b200.device import ops
from kubetorch_b200.serving import codec
from oracle import ref_dispatch as R

json_leaf = st.one_of(st.none(), st.booleans(), st.integers(-2**40, 2**40), st.floats(allow_nan=False, allow_infinity=False),
                      st.text(max_size=12))
json_value = st.recursive(json_leaf, lambda c: st.one_of(st.lists(c, max_size=4), st.dictionaries(st.text(max_size=6), c, max_size=4)),
                          max_leaves=12)


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 10**9), st.integers(1, 64))
def test_shard_bounds_partition_the_range_like_torch_chunk(n, world):
    prev_end, total = 0, 0
    chunk = -(-n // world)
    for r in range(world):
        b, e = ops.shard_bounds(n, world, r)
        cb, ce = ctypes.c_size_t(), ctypes.c_size_t()
        L.call("ktb_shard_bounds", n, world, r, ctypes.byref(cb), ctypes.byref(ce))
        assert (b, e) == (cb.value, ce.value)
        assert b == min(n, prev_end if r else 0) and b <= e <= n and e - b <= chunk
        prev_end, total = e, total + (e - b)
    assert total == n
    if 0 < n <= 4096:
        assert [c.numel() for c in torch.arange(n).chunk(world)] == [
            e - b for b, e in (ops.shard_bounds(n, world, r) for r in range(world)) if e > b]


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(0, 10**7), min_size=0, max_size=40))
def test_pack_layout_offsets_are_aligned_disjoint_and_ordered(sizes):
    offs, total = ops.pack_layout(sizes)
    end = 0
    for o, s in zip(offs, sizes):
        assert o % L.PACK_ALIGN == 0 and o >= end
        end = o + s
    assert total >= end and total % L.PACK_ALIGN == 0 and total - end < L.PACK_ALIGN + (0 if not sizes else L.PACK_ALIGN)


@settings(max_examples=100, deadline=None)
@given(st.lists(json_value, max_size=3), st.dictionaries(st.text(min_size=1, max_size=6), json_value, max_size=3),
       st.sampled_from(["json", "pickle"]))
def test_codecs_agree_with_the_oracle_and_round_trip(args, kwargs, serialization):
    kwargs = {k: v for k, v in kwargs.items() if k not in ("workers", "restart_procs")}
    mine = codec.serialize_body({"args": list(args), "kwargs": dict(kwargs, workers=[0])}, serialization)
    theirs = R.serialize_body(R.build_call_body(*args, **dict(kwargs, workers=[0])), serialization)
    assert mine == theirs and mine["workers"] == [0]                      # byte-identical wire bodies
    wire = json.loads(json.dumps(mine))
    a1, k1 = codec.parse_callable_params(dict(wire), serialization)
    a2, k2 = R.parse_callable_params(dict(wire), serialization)
    assert a1 == a2 == list(args) and k1 == k2 == kwargs
    res = codec.serialize_result(args, serialization)
    assert res == R.serialize_result(args, serialization)
    assert codec.deserialize_response(json.loads(json.dumps([res, res])), serialization) == [list(args)] * 2


@settings(max_examples=60, deadline=None)
@given(st.sampled_from([TypeError, ValueError, KeyError, AssertionError, FileNotFoundError, PermissionError,
                        NotImplementedError, MemoryError, OSError, RuntimeError, ZeroDivisionError]), st.text(max_size=20))
def test_error_envelope_and_status_map_agree_with_the_oracle(exc_type, msg):
    try:
        raise exc_type(msg)
    except Exception as e:  # noqa: BLE001
        env = codec.package_exception(e, pod_name="p")
        status, ref_env = R.package_exception(e)
    assert env["status_code"] == status == R.status_code_for(exc_type(msg))
    assert (env["error_type"], env["message"]) == (ref_env["error_type"], ref_env["message"])
    back = codec.rebuild_exception(env)
    assert isinstance(back, exc_type) and back.pod_name == "p" and "Traceback" in back.remote_traceback


worker_item = st.one_of(st.integers(-2, 6), st.sampled_from(["0", "1", "2", "10", "10.0.0.1", "10.0.0.2", "10.0.0.9", "x"]),
                        st.floats(0, 3))


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 4), st.one_of(st.none(), st.sampled_from(["any", "ready", "0.2", "10.0.0", "zzz"]),
                                    st.lists(worker_item, min_size=1, max_size=4)))
def test_worker_selection_agrees_with_the_oracle(n_nodes, workers_arg):
    """Product: node indices that take part (serving/supervisors.py). Oracle: (remote ips, call_local) of the
    coordinator (spmd_supervisor.py:219-261). Same participants, same error text."""
    from kubetorch_b200.serving.supervisors import select_worker_nodes

    ips = [f"10.0.0.{i + 1}" for i in range(n_nodes)]
    try:
        remote, local = R.select_workers(workers_arg, ips, ips[0])
        want = sorted(([0] if local else []) + [ips.index(ip) for ip in remote])
        err = None
    except ValueError as e:
        want, err = None, str(e)
    if err is not None:
        with pytest.raises(ValueError) as ei:
            select_worker_nodes(workers_arg, ips, ips[0])
        assert str(ei.value) == err
    else:
        assert select_worker_nodes(workers_arg, ips, ips[0]) == want


tensor_dtypes = st.sampled_from([torch.float32, torch.bfloat16, torch.float16, torch.int64, torch.int32, torch.uint8,
                                 torch.bool, torch.float64])


@settings(max_examples=120, deadline=None)
@given(tensor_dtypes, st.lists(st.integers(0, 7), min_size=0, max_size=3), st.integers(0, 3), st.booleans())
def test_fastpickle_equals_stock_pickle_on_tensors(dtype, shape, offset, transpose):
    """The pipe codec must hand the rank exactly what stock pickle would (dtype, shape, values), for views too."""
    import pickle

    from kubetorch_b200.serving import fastpickle as F

    n = 1
    for d in shape:
        n *= d
    base = (torch.arange(n + offset) % 5).to(dtype)
    t = base[offset:].reshape(shape)
    if transpose and t.dim() >= 2:
        t = t.transpose(0, 1)
    payload = {"x": [t, 3], "y": ("s", t[:0] if t.dim() else t)}
    a = F.loads(F.dumps(payload))
    b = pickle.loads(pickle.dumps(payload))
    for got, want in ((a["x"][0], b["x"][0]), (a["y"][1], b["y"][1])):
        assert got.dtype == want.dtype and got.shape == want.shape and torch.equal(got, want)
    assert a["x"][1] == 3 and a["y"][0] == "s"

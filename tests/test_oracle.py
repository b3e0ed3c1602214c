"""not-gpu: the oracle restatement (oracle/ref_dispatch.py) against (a) the reference's own golden
vectors and (b) outputs recorded from the UNMODIFIED reference runtime (tests/golden/ref_runtime.pt)."""
import json
import os

import pytest
import torch

from conftest import REPO, resolve_args
from oracle import cases, ref_dispatch as R


def _base_name(exc):
    """Class name of a rehydrated exception (the reference raises RemoteException(<original class>))."""
    return next(c.__name__ for c in type(exc).__mro__ if c.__name__ != "RemoteException")


def _callable(rec_or_case):
    return getattr(cases, rec_or_case["callable"])


def _run_oracle(rec, args):
    cfg = rec["distributed_config"]
    fn = _callable(rec)
    ser = rec.get("serialization", "json")
    kwargs = dict(rec.get("kwargs") or {})
    if rec.get("method"):
        inst = fn()
        target = getattr(inst, rec["method"])
    else:
        target = fn
    if cfg["distribution_type"] == "local":
        return R.local_call(target, *args, serialization=ser, allowed=rec.get("allowed"), **kwargs)
    if rec.get("pods"):   # recorded on several real pods (uvicorn on 127.0.0.k)
        return R.multipod_call(target, *args, num_proc=cfg["num_proc"], pod_ips=rec["pods"],
                               distribution_type=cfg["distribution_type"], serialization=ser,
                               allowed=rec.get("allowed"), **kwargs)
    return R.spmd_call(target, *args, num_proc=cfg["num_proc"], distribution_type=cfg["distribution_type"],
                       serialization=ser, allowed=rec.get("allowed"), **kwargs)


def test_reference_asset_goldens():
    with open(os.path.join(REPO, "tests", "golden", "reference_assets.json")) as f:
        assets = json.load(f)
    for case in assets["cases"]:
        if case["callable"] == "torch_ddp":
            continue  # needs real rank processes; covered in test_api_cpu.py
        if "expected" in case:
            assert _run_oracle(case, case["args"]) == case["expected"], case["name"]
        else:
            with pytest.raises(Exception) as ei:
                _run_oracle(case, case["args"])
            assert _base_name(ei.value) == case["error"], case["name"]
            assert R.status_code_for(getattr(__import__("builtins"), case["error"])()) == case["error_code"]


def _same(a, b):
    if isinstance(a, torch.Tensor):
        return isinstance(b, torch.Tensor) and a.dtype == b.dtype and a.shape == b.shape and \
            torch.equal(a.contiguous().reshape(-1).view(torch.uint8), b.contiguous().reshape(-1).view(torch.uint8))
    if isinstance(a, (list, tuple)):
        return isinstance(b, (list, tuple)) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, dict):
        return isinstance(b, dict) and a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    return a == b


def test_oracle_matches_recorded_reference_runtime(golden):
    """Every recorded call: same result (bit-exact tensors) or same error type / message / status."""
    n = 0
    for name, rec in golden["cases"].items():
        args = resolve_args(golden, rec["args"])
        if name == "number_count" or name.startswith(("number_state_", "torch_ddp_", "all_reduce_", "mp_all_reduce_")):
            continue  # call history of one deployment / real torch.distributed ranks: replayed in test_api_cpu.py
        if rec["status_code"] == 200:
            got = _run_oracle(rec, args)
            assert _same(got, rec["result"]), name
        else:
            with pytest.raises(Exception) as ei:
                _run_oracle(rec, args)
            err = rec["error"]
            assert _base_name(ei.value) == err["error_type"], name
            assert ei.value.args[0].split("\n\n")[0] == err["message"], name
            assert ei.value.http_status == rec["status_code"], name
        n += 1
    assert n >= 30


def test_wire_codec_is_byte_compatible_with_reference(golden):
    """serialize_body / deserialize_response produce and accept the reference's wire format."""
    x = golden["all_inputs"]["f32_1003"]
    body = R.serialize_body(R.build_call_body(x, workers=[0], restart_procs=False, k=1), "pickle")
    assert set(body) == {"data", "workers", "restart_procs"}  # magic kwargs hoisted (serving/utils.py:739-741)
    import base64
    import pickle

    payload = pickle.loads(base64.b64decode(body["data"]))
    assert list(payload) == ["args", "kwargs"] and payload["kwargs"] == {"k": 1}
    assert torch.equal(payload["args"][0], x)
    args, kwargs = R.parse_callable_params(json.loads(json.dumps(body)), "pickle")
    assert torch.equal(args[0], x) and kwargs == {"k": 1}
    env = R.serialize_result([x, 3], "pickle")
    assert _same(R.deserialize_response(json.loads(json.dumps([env, env])), "pickle"), [[x, 3], [x, 3]])
    with pytest.raises(R.SerializationError):
        R.serialize_result({1, 2}, "json")


def test_env_contract_and_selectors():
    env = R.pytorch_env(["10.0.0.1", "10.0.0.2"], node_rank=1, local_rank=2, num_local_procs=4, port=None)
    assert env == {"WORLD_SIZE": "8", "RANK": "6", "LOCAL_RANK": "2", "NODE_RANK": "1",
                   "POD_IPS": "10.0.0.1,10.0.0.2", "MASTER_ADDR": "10.0.0.1", "MASTER_PORT": "12345"}
    ips = ["10.0.0.1", "10.0.0.2", "10.0.0.3"]
    assert R.select_workers(None, ips, ips[0]) == (ips[1:], True)
    assert R.select_workers("any", ips, ips[0]) == ([], True)
    assert R.select_workers([1, "2"], ips, ips[0]) == (ips[1:], False)
    assert R.select_workers(["10.0.0.1"], ips, ips[0]) == ([], True)
    assert R.select_workers("0.3", ips, ips[0]) == ([ips[2]], True)
    with pytest.raises(ValueError, match="Worker index 10 out of range. Valid range: 0-2"):
        R.select_workers([10], ips, ips[0])
    with pytest.raises(ValueError, match="Invalid worker specification"):
        R.select_workers([1.5], ips, ips[0])


def test_exception_rehydration_shape():
    try:
        raise KeyError("missing")
    except KeyError as e:
        status, env = R.package_exception(e)
    assert status == 404 and env["error_type"] == "KeyError"
    exc = R.rehydrate_exception(env)
    assert isinstance(exc, KeyError) and exc.pod_name == env["pod_name"] and "Traceback" in exc.remote_traceback
    assert "Traceback" in str(exc)
    dyn = R.rehydrate_exception({"error_type": "WeirdError", "message": "m", "traceback": "tb", "pod_name": "p"})
    assert _base_name(dyn) == "WeirdError"


def test_oracle_runtime_with_real_processes():
    x = torch.arange(1003, dtype=torch.float32)
    with R.OracleRuntime("oracle.cases", "double", 3, "spmd", extra_path=REPO) as rt:
        out = rt.call(x, serialization="pickle")
    assert [o.numel() for o in out] == [335, 335, 333] and torch.equal(torch.cat(out), x * 2)


def test_tree_fanout_matches_reference_outputs(golden):
    """get_tree_children (spmd_supervisor.py:68-101) restated; vectors produced by calling the reference method."""
    assert len(golden["tree_children"]) >= 20
    for rec in golden["tree_children"]:
        ips = sorted(f"10.0.{i // 250}.{i % 250}" for i in range(rec["n"]))
        assert R.tree_children(ips, rec["ip"], rec["fanout"]) == rec["children"], rec
    ips = [f"10.1.0.{i}" for i in range(5)]
    assert R.fanout_targets(ips, ips[0]) == ips[1:]                      # flat below 100 pods
    big = sorted(f"10.2.{i // 200}.{i % 200}" for i in range(150))
    assert R.fanout_targets(big, big[0]) == big[1:51] and R.fanout_targets(big, big[1]) == big[51:101]
    assert R.fanout_targets(big, big[2]) == big[101:150] and R.fanout_targets(big, big[3]) == []

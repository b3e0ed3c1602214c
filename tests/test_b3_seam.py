"""not-gpu, authoring container only: B200Supervisor behind the REAL reference server.

The unmodified reference (`/root/reference/python_client`) runs its FastAPI app under TestClient with the 3-line
supervisor_factory hook of INTEGRATION.md applied; requests are built and decoded with the reference's own client
codecs.  This pins seam B3: construction from KT_DISTRIBUTED_CONFIG alone, callable from the KT_*
environment, raw `{"data": b64}` bodies in, per-rank `{"data": b64}` out, `workers=` errors, allow-list errors.
The device layer is a torch-CPU stub here (no GPU in this container); tests/test_gpu_api.py::test_b3_* run the same
contract against the real kernels with recorded reference requests."""
import json
import os
import subprocess
import sys
import tempfile

import pytest
import torch

from conftest import REPO

REFERENCE = "/root/reference/python_client"
pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is only mounted in the authoring container")

STUB = (
    "class WebSocketException(Exception): pass\n"
    "ABNF = type('ABNF', (), {'OPCODE_BINARY': 2, 'OPCODE_TEXT': 1})\n"
    "def create_connection(*a, **k): raise RuntimeError('stub')\n"
)


def _run(cfg):
    work = tempfile.mkdtemp(prefix="kt_b3_")
    with open(os.path.join(work, "websocket.py"), "w") as f:   # the one absent import of the reference (never used here)
        f.write(STUB)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([work, REFERENCE, REPO])
    env["HOME"] = work
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    cfg = dict(cfg, repo=REPO)
    p = subprocess.run([sys.executable, os.path.join(REPO, "tests", "b3_driver.py"), "--stub", json.dumps(cfg)],
                       env=env, cwd=work, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("B3RESULT ")]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-4000:]
    return json.loads(lines[-1][len("B3RESULT "):])


def _t(dtype, shape, seed=0):
    return {"tensor": {"dtype": dtype, "shape": shape, "seed": seed}}


def _make(spec):
    g = torch.Generator().manual_seed(spec["tensor"].get("seed", 0))
    dt = getattr(torch, spec["tensor"]["dtype"])
    if dt.is_floating_point:
        return torch.randn(spec["tensor"]["shape"], generator=g).to(dt)
    return torch.randint(-1000, 1000, spec["tensor"]["shape"], generator=g, dtype=dt)


def _shards(x, world):
    ch = x.chunk(world)
    return [ch[r] if r < len(ch) else x[:0] for r in range(world)]


def test_reference_server_drives_b200_supervisor_with_raw_pickle_bodies():
    x, small = _t("float32", [1003]), _t("float32", [3])
    out = _run({"callable": "double", "distributed_config": {"distribution_type": "b200", "num_proc": 4, "self_check": False},
                "calls": [{"args": [x]}, {"args": [small]}, {"args": [_t("float32", [10, 7], 3)]},
                          {"args": [x], "kwargs": {"workers": [10]}}, {"args": [x], "kwargs": {"workers": [1.5]}},
                          {"args": [x], "kwargs": {"workers": "any"}}]})
    recs = out["records"]
    for rec, spec in zip(recs[:3], (x, small, _t("float32", [10, 7], 3))):
        assert rec["status_code"] == 200, rec
        want = [s * 2 for s in _shards(_make(spec), 4)]
        assert len(rec["result"]) == 4
        for got, w in zip(rec["result"], want):
            assert got["dtype"] == str(w.dtype) and got["shape"] == list(w.shape)
            assert got["data"] == w.reshape(-1).tolist()
    # the reference's selector errors, through the reference's own exception handler (recorded: workers_bad_index/spec)
    assert recs[3]["status_code"] == 400 and recs[3]["error"]["error_type"] == "ValueError"
    assert recs[3]["error"]["message"] == "Worker index 10 out of range. Valid range: 0-0"
    assert recs[4]["status_code"] == 400 and recs[4]["error"]["message"] == (
        "Invalid worker specification: 1.5. Must be an IP address, integer index, or numeric string.")
    assert recs[5]["status_code"] == 200 and len(recs[5]["result"]) == 4      # "any" = the coordinator's local ranks
    assert ("map_host_multi", 4) in [tuple(c) for c in out["device_calls"]]


def test_reference_server_multi_node_config_and_workers_subset():
    """quorum_workers=2 x num_proc=2 (the recorded mp_double_* shape): `workers=[1]` returns node 1's ranks only,
    which keep their GLOBAL rank / world size (recorded shapes [(251,), (250,)])."""
    x = _t("float32", [1003])
    out = _run({"callable": "double",
                "distributed_config": {"distribution_type": "b200", "num_proc": 2, "quorum_workers": 2, "self_check": False},
                "calls": [{"args": [x]}, {"args": [x], "kwargs": {"workers": [1]}}, {"args": [x], "kwargs": {"workers": [7]}}]})
    full, sub, bad = out["records"]
    want = [s * 2 for s in _shards(_make(x), 4)]
    assert [r["shape"] for r in full["result"]] == [[251], [251], [251], [250]]
    assert [r["shape"] for r in sub["result"]] == [[251], [250]]
    assert sub["result"][0]["data"] == want[2].tolist() and sub["result"][1]["data"] == want[3].tolist()
    assert bad["status_code"] == 400 and bad["error"]["message"] == "Worker index 7 out of range. Valid range: 0-1"


def test_reference_server_allow_list_json_mode_and_class_callable():
    out = _run({"callable": "double", "allowed": "json",
                "distributed_config": {"distribution_type": "b200", "num_proc": 2, "self_check": False},
                "calls": [{"args": [_t("float32", [8])], "serialization": "pickle"}]})
    rec = out["records"][0]
    assert rec["status_code"] == 400
    assert "Serialization format 'pickle' not allowed. Allowed formats: ['json']" in json.dumps(rec["error"])
    # a class: instance built from KT_INIT_ARGS, method from the URL, kwargs-bound parameters of an affine op
    out = _run({"callable": "Scaler", "init_args": {"tag": "t"},
                "distributed_config": {"distribution_type": "b200", "num_proc": 3, "self_check": False},
                "calls": [{"method": "triple", "args": [_t("int64", [130])]}, {"method": "nope", "args": []}]})
    ok, missing = out["records"]
    want = [s * 3 for s in _shards(_make(_t("int64", [130])), 3)]
    assert [r["data"] for r in ok["result"]] == [w.tolist() for w in want]
    assert missing["status_code"] == 404 and "Method 'nope' not found in class 'Scaler'" in json.dumps(missing["error"])
    out = _run({"callable": "affine", "distributed_config": {"distribution_type": "b200", "num_proc": 2, "self_check": False},
                "calls": [{"args": [_t("float32", [1001], 5), 0.1], "kwargs": {"beta": 0.3}},
                          {"args": [_t("int64", [130]), 0.5, 1]}]})
    rec, bad = out["records"]
    want = [s * 0.1 + 0.3 for s in _shards(_make(_t("float32", [1001], 5)), 2)]
    assert [r["data"] for r in rec["result"]] == [w.tolist() for w in want]
    assert bad["status_code"] == 422 and bad["error"]["error_type"] == "TypeError"   # non-integral alpha on an int tensor

"""Generate tests/golden/ref_runtime.pt by running the UNMODIFIED reference pod runtime.

TEST INFRASTRUCTURE.  Runs only in the authoring container, where /root/reference exists:
imports kubetorch 0.5.0 from /root/reference/python_client (read-only; nothing is copied),
drives its FastAPI app in-process with fastapi.testclient.TestClient — lifespan → load_callable →
supervisor_factory → ProcessPool → spawned ProcessWorkers — exactly as the reference's own
tests/test_http_server.py:86-96 does, and records request → response pairs for the callables in
oracle/cases.py.  The only shim is a 3-line stub for the absent `websocket-client` package
(imported at kt/data_store/websocket_tunnel.py:8, never used on this path), written to a temp dir.

Usage:  python oracle/make_golden.py            # regenerates tests/golden/ref_runtime.pt
"""
from __future__ import annotations

import json
import os
import pickle
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference/python_client"
OUT = os.path.join(REPO, "tests", "golden", "ref_runtime.pt")

STUB = (
    "class WebSocketException(Exception): pass\n"
    "ABNF = type('ABNF', (), {'OPCODE_BINARY': 2, 'OPCODE_TEXT': 1})\n"
    "def create_connection(*a, **k): raise RuntimeError('stub')\n"
)


from oracle.golden_inputs import make_inputs as _inputs, tensor_sha256  # noqa: E402


# group → (callable name, distributed config, allowed serialization, [(case, method, args-spec, kwargs, serialization)])
# args-spec entries that are strings starting with "@" name a tensor from _inputs().
GROUPS = {
    "summer_local": ("summer", {"distribution_type": "local"}, "json,pickle", [
        ("summer_valid", None, [1, 2], {}, "json"),
        ("summer_invalid", None, ["a", 2], {}, "json"),
        ("summer_pickle", None, [1, 2], {}, "pickle"),
    ]),
    "hello_local": ("hello_world", {"distribution_type": "local"}, "json,pickle", [
        ("hello_world", None, [], {}, "json"),
    ]),
    "number_local": ("Number", {"distribution_type": "local"}, "json,pickle", [
        ("number_add", "add", [1, 2], {}, "json"),
        ("number_add_invalid", "add", ["a", 2], {}, "json"),
        ("number_count", "count", [], {}, "json"),
    ]),
    "torch_summer_pt4": ("torch_summer", {"distribution_type": "pytorch", "num_proc": 4}, "json,pickle", [
        ("torch_summer_valid", None, [1, 2], {}, "json"),
        ("torch_summer_invalid", None, ["a", 2], {}, "json"),
    ]),
    "env_pt4": ("env_report", {"distribution_type": "pytorch", "num_proc": 4}, "json,pickle", [
        ("env_pytorch_4", None, [], {}, "json"),
        ("workers_bad_index", None, [], {"workers": [10]}, "json"),
        ("workers_bad_spec", None, [], {"workers": [1.5]}, "json"),
        ("workers_any", None, [], {"workers": "any"}, "json"),
    ]),
    "env_spmd2": ("env_report", {"distribution_type": "spmd", "num_proc": 2}, "json,pickle", [
        ("env_spmd_2", None, [], {}, "json"),
    ]),
    "raise_spmd2": ("raise_value_error", {"distribution_type": "spmd", "num_proc": 2}, "json,pickle", [
        ("raise_value_error", None, ["boom"], {}, "json"),
    ]),
    "pickle_forbidden": ("summer", {"distribution_type": "local"}, "json", [
        ("pickle_not_allowed", None, [1, 2], {}, "pickle"),
    ]),
    "double_spmd4": ("double", {"distribution_type": "spmd", "num_proc": 4}, "json,pickle", [
        ("double_f32_1003_x4", None, ["@f32_1003"], {}, "pickle"),
        ("double_f32_3_x4", None, ["@f32_3"], {}, "pickle"),  # fewer elements than ranks
        ("double_bf16_777_x4", None, ["@bf16_777"], {}, "pickle"),
    ]),
    "identity_spmd3": ("identity", {"distribution_type": "spmd", "num_proc": 3}, "json,pickle", [
        ("identity_u8_1000_x3", None, ["@u8_1000"], {}, "pickle"),
        ("identity_i64_130_x3", None, ["@i64_130"], {}, "pickle"),
    ]),
    "affine_spmd2": ("affine", {"distribution_type": "spmd", "num_proc": 2}, "json,pickle", [
        ("affine_bf16_777_x2", None, ["@bf16_777", 1.5, 0.25], {}, "pickle"),
        ("affine_bf16_777_x2_inexact_scalars", None, ["@bf16_777", 1.7, -0.3], {}, "pickle"),
        ("affine_f32_1001_x2", None, ["@f32_rand_1001", 0.1, 0.3], {}, "pickle"),
        ("affine_i32_515_x2", None, ["@i32_515", 3, -7], {}, "pickle"),
        ("affine_i64_130_x2", None, ["@i64_130", -5, 11], {}, "pickle"),
    ]),
    "affine_f16_spmd3": ("affine", {"distribution_type": "spmd", "num_proc": 3}, "json,pickle", [
        ("affine_f16_515_x3", None, ["@f16_515", 1.7, -0.3], {}, "pickle"),
    ]),
    "scale_spmd4": ("scale", {"distribution_type": "spmd", "num_proc": 4}, "json,pickle", [
        ("scale_f32_1001_x4", None, ["@f32_rand_1001", 0.1], {}, "pickle"),
        ("scale_i32_515_x4", None, ["@i32_515", 65537], {}, "pickle"),  # wraps
    ]),
    "sum_spmd4": ("shard_sum", {"distribution_type": "spmd", "num_proc": 4}, "json,pickle", [
        ("sum_i64_130_x4", None, ["@i64_130"], {}, "pickle"),
        ("sum_i32_515_x4", None, ["@i32_515", 3, 1], {}, "pickle"),
        ("sum_f32_1001_x4", None, ["@f32_rand_1001"], {}, "pickle"),
    ]),
    "spmd_identity_2": ("spmd_identity", {"distribution_type": "spmd", "num_proc": 2}, "json,pickle", [
        ("broadcast_f32_3_x2", None, ["@f32_3"], {}, "pickle"),
    ]),
    "mixed_spmd2": ("mixed_payload", {"distribution_type": "spmd", "num_proc": 2}, "json,pickle", [
        ("mixed_payload_x2", None, ["@f32_3", {"t": "@i64_130", "tag": "hello"}], {"scale": 2}, "pickle"),
    ]),
    "number_spmd2": ("Number", {"distribution_type": "spmd", "num_proc": 2}, "json,pickle", [
        ("number_spmd_add", "add", [1, 2], {}, "json"),
        ("number_spmd_add_kwargs", "add", [], {"x": 5, "y": 6}, "json"),
    ]),
    "async_local": ("async_summer", {"distribution_type": "local"}, "json,pickle", [
        ("async_summer_1", None, [1, 2], {"sleep_time": 0.01}, "json"),
        ("async_summer_2", None, [10, 20], {"sleep_time": 0.01}, "json"),
    ]),
    # real torch.distributed (gloo) rank processes brought up from the env contract by the reference launcher
    "torch_ddp_pt4": ("torch_ddp", {"distribution_type": "pytorch", "num_proc": 4}, "json,pickle", [
        ("torch_ddp_valid_recorded", None, [3], {}, "json"),
        ("torch_ddp_invalid_recorded", None, ["a"], {}, "json"),
    ]),
    "all_reduce_pt4": ("all_reduce_rank", {"distribution_type": "pytorch", "num_proc": 4}, "json,pickle", [
        ("all_reduce_rank_pt4", None, [], {}, "json"),
    ]),
    # per-rank state of a class persists across calls (ordered sequence on one deployment)
    "number_state_spmd2": ("Number", {"distribution_type": "spmd", "num_proc": 2}, "json,pickle", [
        ("number_state_0_count", "count", [], {}, "json"),
        ("number_state_1_add", "add", [2, 3], {}, "json"),
        ("number_state_2_add", "add", [4, 5], {}, "json"),
        ("number_state_3_count", "count", [], {}, "json"),
    ]),
    "mlp_spmd2": ("mlp_policy", {"distribution_type": "spmd", "num_proc": 2}, "json,pickle", [
        ("mlp_bf16_256_x2", None, ["@mlp_obs", "@mlp_w1", "@mlp_w2", "@mlp_w3"], {}, "pickle"),
    ]),
}


# Two real pods (uvicorn on 127.0.0.1 / 127.0.0.2, the recipe of SURVEY.md §8(c)) x 2 ranks each: pins the cross-pod
# half of the path (kt/serving/remote_worker_pool.py:116-510, spmd_supervisor.py:219-261 `workers=` selectors,
# rank ordering across pods).  group -> (callable, distributed config, n_pods, cases)
MULTIPOD_GROUPS = {
    "mp_env_2x2": ("env_report", {"distribution_type": "pytorch", "num_proc": 2, "quorum_workers": 2}, 2, [
        ("mp_env_all", None, [], {}, "json"),
        ("mp_env_workers_0", None, [], {"workers": [0]}, "json"),
        ("mp_env_workers_str1", None, [], {"workers": ["1"]}, "json"),
        ("mp_env_workers_0_1", None, [], {"workers": [0, 1]}, "json"),
        ("mp_env_workers_any", None, [], {"workers": "any"}, "json"),
        ("mp_env_workers_bad", None, [], {"workers": [10]}, "json"),
    ]),
    "mp_raise_2x2": ("raise_value_error", {"distribution_type": "spmd", "num_proc": 2, "quorum_workers": 2}, 2, [
        ("mp_raise_value_error", None, ["boom across pods"], {}, "json"),
    ]),
    "mp_all_reduce_2x2": ("all_reduce_rank", {"distribution_type": "pytorch", "num_proc": 2, "quorum_workers": 2}, 2, [
        ("mp_all_reduce_rank_2x2", None, [], {}, "json"),     # gloo group spanning both pods: 0+1+2+3
    ]),
    "mp_double_2x2": ("double", {"distribution_type": "spmd", "num_proc": 2, "quorum_workers": 2}, 2, [
        ("mp_double_f32_1003_2x2", None, ["@f32_1003"], {}, "pickle"),
        ("mp_double_f32_1003_workers_1", None, ["@f32_1003"], {"workers": [1]}, "pickle"),
    ]),
}


def run_multipod_group(group: str, out_path: str):
    """Child process: start one uvicorn pod per IP, POST to pod 0 as the client would, record, stop the pods."""
    import socket
    import time

    import httpx
    from kubetorch.resources.callables.utils import build_call_body
    from kubetorch.serving.utils import _deserialize_response, _serialize_body

    name, dist_cfg, n_pods, cases = MULTIPOD_GROUPS[group]
    ips = [f"127.0.0.{k + 1}" for k in range(n_pods)]
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    base = dict(os.environ)
    base.update({
        "KT_LOG_STREAMING_ENABLED": "false", "KT_METRICS_ENABLED": "false", "POD_NAMESPACE": "kubetorch",
        "LOCAL_IPS": ",".join(ips), "KT_SERVICE_NAME": "golden", "KT_FILE_PATH": REPO, "KT_MODULE_NAME": "oracle.cases",
        "KT_CLS_OR_FN_NAME": name, "KT_INIT_ARGS": "null", "KT_ALLOWED_SERIALIZATION": "json,pickle",
        "KT_DISTRIBUTED_CONFIG": json.dumps(dist_cfg), "KT_SERVER_PORT": str(port),
    })
    pods = []
    try:
        for k, ip in enumerate(ips):
            env = dict(base)
            env.update({"POD_IP": ip, "POD_NAME": f"golden-pod-{k}"})
            pods.append(subprocess.Popen(
                [sys.executable, "-m", "uvicorn", "kubetorch.serving.http_server:app", "--host", ip, "--port", str(port),
                 "--log-level", "warning"], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        deadline = time.time() + 180
        for ip in ips:                      # every pod answers its health route before the first call
            while True:
                try:
                    if httpx.get(f"http://{ip}:{port}/health", timeout=2).status_code == 200:
                        break
                except Exception:  # noqa: BLE001
                    pass
                if time.time() > deadline:
                    raise RuntimeError(f"pod {ip} did not come up")
                time.sleep(0.5)
        inputs = _inputs()
        results = {}
        for case, method, arg_spec, kwargs, ser in cases:
            args = _resolve(arg_spec, inputs)
            body = _serialize_body(build_call_body(*args, **dict(kwargs)), ser)
            url = f"http://{ips[0]}:{port}/{name}" + (f"/{method}" if method else "")
            resp = httpx.post(url, json=body, headers={"X-Serialization": ser, "X-Request-ID": case}, timeout=300)
            rec = {"callable": name, "method": method, "distributed_config": dist_cfg, "allowed": "json,pickle",
                   "args": arg_spec, "kwargs": kwargs, "serialization": ser, "status_code": resp.status_code,
                   "pods": ips}
            if resp.status_code == 200:
                rec["result"] = _deserialize_response(resp, ser)
            else:
                err = resp.json()
                rec["error"] = {k: err.get(k) for k in ("error_type", "message", "pod_name", "detail") if k in err}
            results[case] = rec
        with open(out_path, "wb") as f:
            pickle.dump(results, f)
    finally:
        for p in pods:
            p.terminate()
        for p in pods:
            try:
                p.wait(timeout=20)
            except subprocess.TimeoutExpired:
                p.kill()


def _resolve(spec, inputs):
    """Replace "@name" strings (at any depth of lists/dicts) by the named input tensor."""
    if isinstance(spec, str) and spec.startswith("@"):
        return inputs[spec[1:]]
    if isinstance(spec, list):
        return [_resolve(v, inputs) for v in spec]
    if isinstance(spec, dict):
        return {k: _resolve(v, inputs) for k, v in spec.items()}
    return spec


def run_group(group: str, out_path: str):
    """Child process: one callable deployment (env-var "metadata"), several calls."""
    name, dist_cfg, allowed, cases = GROUPS[group]
    os.environ["KT_LOG_STREAMING_ENABLED"] = "false"
    os.environ["KT_METRICS_ENABLED"] = "false"
    os.environ.update({
        "POD_NAMESPACE": "kubetorch", "POD_NAME": "golden-pod", "POD_IP": "localhost", "LOCAL_IPS": "localhost",
        "KT_SERVICE_NAME": "golden", "KT_FILE_PATH": REPO, "KT_MODULE_NAME": "oracle.cases",
        "KT_CLS_OR_FN_NAME": name, "KT_INIT_ARGS": "null", "KT_ALLOWED_SERIALIZATION": allowed,
        "KT_DISTRIBUTED_CONFIG": json.dumps(dist_cfg),
    })
    if name == "Number":
        os.environ["KT_CLS_OR_FN_NAME"] = "Number"
    from fastapi.testclient import TestClient
    from kubetorch.resources.callables.utils import build_call_body
    from kubetorch.serving.http_server import app
    from kubetorch.serving.utils import _deserialize_response, _serialize_body

    inputs = _inputs()
    results = {}
    with TestClient(app, raise_server_exceptions=False) as client:
        for case, method, arg_spec, kwargs, ser in cases:
            args = _resolve(arg_spec, inputs)
            body = _serialize_body(build_call_body(*args, **dict(kwargs)), ser)
            url = f"/{name}/{method}" if method else f"/{name}"
            resp = client.post(url, json=body, headers={"X-Serialization": ser, "X-Request-ID": case})
            rec = {
                "callable": name, "method": method, "distributed_config": dist_cfg, "allowed": allowed,
                "args": arg_spec, "kwargs": kwargs, "serialization": ser, "status_code": resp.status_code,
            }
            if resp.status_code == 200:
                rec["result"] = _deserialize_response(resp, ser)
            else:
                err = resp.json()
                rec["error"] = {k: err.get(k) for k in ("error_type", "message", "pod_name", "detail") if k in err}
            results[case] = rec
    with open(out_path, "wb") as f:
        pickle.dump(results, f)


def main():
    import torch

    if len(sys.argv) == 4 and sys.argv[1] == "--group":
        (run_multipod_group if sys.argv[2] in MULTIPOD_GROUPS else run_group)(sys.argv[2], sys.argv[3])
        return
    if not os.path.isdir(REFERENCE):
        raise SystemExit(f"{REFERENCE} not found: goldens can only be regenerated where the reference is mounted")
    work = tempfile.mkdtemp(prefix="kt_golden_")
    with open(os.path.join(work, "websocket.py"), "w") as f:
        f.write(STUB)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([work, REFERENCE, REPO])
    env["HOME"] = work
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    all_results = {}
    ALL = {**GROUPS, **MULTIPOD_GROUPS}
    only = [a for a in sys.argv[1:] if a in ALL or a == "--helpers-only"]
    if only and os.path.exists(OUT):
        all_results.update(torch.load(OUT, weights_only=False)["cases"])   # keep the other groups' records
    for group in ([g for g in only if g in ALL] if only else ALL):
        out_path = os.path.join(work, f"{group}.pkl")
        print(f"[make_golden] {group} ...", flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--group", group, out_path], env=env, check=True,
                       cwd=work, timeout=600)
        with open(out_path, "rb") as f:
            all_results.update(pickle.load(f))
    # pure helper of the path: the fan-out tree (no runtime needed; called on the reference class directly)
    sys.path[:0] = [work, REFERENCE]
    os.environ.setdefault("KT_LOG_STREAMING_ENABLED", "false")
    os.environ.setdefault("KT_METRICS_ENABLED", "false")
    from kubetorch.serving.spmd.spmd_supervisor import SPMDDistributedSupervisor as _RefSup

    tree = []
    for n, fan in ((1, 50), (7, 2), (120, 50), (300, 50), (1000, 10)):
        ips = sorted(f"10.0.{i // 250}.{i % 250}" for i in range(n))
        for who in {ips[0], ips[min(1, n - 1)], ips[n // 2], ips[-1]}:
            tree.append({"n": n, "fanout": fan, "ip": who,
                         "children": _RefSup.get_tree_children(None, ips, who, fan)})
        tree.append({"n": n, "fanout": fan, "ip": "192.168.0.1", "children": _RefSup.get_tree_children(None, ips, "192.168.0.1", fan)})
    fixture = {
        "tree_children": tree,
        "provenance": {
            "generator": "oracle/make_golden.py",
            "reference": "run-house/kubetorch @ 96fac95 (python_client v0.5.0), unmodified, via fastapi TestClient",
            "torch": torch.__version__,
        },
        # small inputs are stored; large ones (MLP weights) only as sha256 — tests regenerate them
        # with oracle.golden_inputs.make_inputs() and verify the hash
        "inputs": {k: v for k, v in _inputs().items() if v.numel() <= 70000},
        "input_sha256": {k: tensor_sha256(v) for k, v in _inputs().items()},
        "cases": all_results,
    }
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(fixture, OUT)
    print(f"[make_golden] wrote {OUT}: {len(all_results)} cases, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()

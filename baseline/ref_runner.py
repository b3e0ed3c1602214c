#!/usr/bin/env python
"""Times the UNMODIFIED reference runtime (baseline/_ref/kubetorch, see install_reference.py) on the host cores.

Two shapes of the reference's dispatch path:
  testclient   1 pod x P ranks: the FastAPI app in-process under fastapi.testclient.TestClient (the reference's own
               tests/test_http_server.py:86-96 recipe) -> SPMDDistributedSupervisor -> ProcessPool -> P spawned
               ProcessWorkers; the client side is the reference's _serialize_body / _deserialize_response.
  pods         N pods x 1 rank: one `uvicorn kubetorch.serving.http_server:app` per pod on 127.0.0.k, real HTTP
               between the coordinator pod and the others (serving/remote_worker_pool.py:254-316) — the closest
               thing to "K8s-pod dispatch" one box offers.
The callable is oracle/cases.py:double (`x.chunk(WORLD_SIZE)[RANK] * 2`), the payload a seeded fp32 tensor.
Prints one JSON object: {"per_call_s": [...], "warm_s": [...], "ok": bool, ...}.

This file runs as a CHILD of bench.py with PYTHONPATH = baseline/_ref : repo.  It is benchmark infrastructure for the
reference arm: nothing under kubetorch_b200/ imports it.
"""
import argparse
import json
import os
import socket
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def base_env(n_ranks_per_pod, extra_cfg=None):
    cfg = {"distribution_type": "spmd", "num_proc": n_ranks_per_pod}
    cfg.update(extra_cfg or {})
    return {
        "KT_LOG_STREAMING_ENABLED": "false", "KT_METRICS_ENABLED": "false", "POD_NAMESPACE": "kubetorch",
        "KT_SERVICE_NAME": "bench", "KT_FILE_PATH": REPO, "KT_MODULE_NAME": "oracle.cases",
        "KT_CLS_OR_FN_NAME": "double", "KT_INIT_ARGS": "null", "KT_ALLOWED_SERIALIZATION": "json,pickle",
        "KT_DISTRIBUTED_CONFIG": json.dumps(cfg),
    }


def run_testclient(args):
    os.environ.update(base_env(args.ranks))
    os.environ.update({"POD_NAME": "bench-pod", "POD_IP": "localhost", "LOCAL_IPS": "localhost"})
    import torch
    from fastapi.testclient import TestClient
    from kubetorch.resources.callables.utils import build_call_body
    from kubetorch.serving.http_server import app
    from kubetorch.serving.utils import _deserialize_response, _serialize_body

    torch.manual_seed(0)
    x = torch.randn(args.elems, dtype=torch.float32)
    warm, per = [], []
    with TestClient(app, raise_server_exceptions=False) as client:
        def call():
            body = _serialize_body(build_call_body(x), "pickle")
            resp = client.post("/double", json=body, headers={"X-Serialization": "pickle"})
            assert resp.status_code == 200, resp.text[:500]
            return _deserialize_response(resp, "pickle")

        ok = True
        for i in range(args.warmup):
            t0 = time.perf_counter()
            out = call()
            warm.append(time.perf_counter() - t0)
            if i == 0:
                ok = len(out) == args.ranks and bool(torch.equal(torch.cat(out), x * 2))
        for _ in range(args.steps):
            t0 = time.perf_counter()
            call()
            per.append(time.perf_counter() - t0)
    print("REFRESULT " + json.dumps({"shape": f"1 pod x {args.ranks} ranks (TestClient)", "per_call_s": per, "warm_s": warm,
                                     "ok": ok, "elems": args.elems, "ranks": args.ranks}), flush=True)


def run_pods(args):
    import httpx
    import torch
    from kubetorch.resources.callables.utils import build_call_body
    from kubetorch.serving.utils import _deserialize_response, _serialize_body

    n = args.ranks
    ips = [f"127.0.0.{k + 1}" for k in range(n)]
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env0 = dict(os.environ)
    env0.update(base_env(1, {"quorum_workers": n, "quorum_timeout": 120}))
    env0.update({"LOCAL_IPS": ",".join(ips), "KT_SERVER_PORT": str(port)})
    pods = []
    try:
        for k, ip in enumerate(ips):
            env = dict(env0)
            env.update({"POD_IP": ip, "POD_NAME": f"bench-pod-{k}"})
            pods.append(subprocess.Popen(
                [sys.executable, "-m", "uvicorn", "kubetorch.serving.http_server:app", "--host", ip, "--port", str(port),
                 "--log-level", "warning"], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                start_new_session=True))
        deadline = time.time() + 240
        for ip in ips:
            while True:
                try:
                    if httpx.get(f"http://{ip}:{port}/health", timeout=2).status_code == 200:
                        break
                except Exception:  # noqa: BLE001
                    pass
                if time.time() > deadline:
                    raise RuntimeError(f"pod {ip} did not come up")
                time.sleep(0.5)
        torch.manual_seed(0)
        x = torch.randn(args.elems, dtype=torch.float32)
        client = httpx.Client(timeout=None)

        def call():
            body = _serialize_body(build_call_body(x), "pickle")
            resp = client.post(f"http://{ips[0]}:{port}/double", json=body,
                               headers={"X-Serialization": "pickle", "X-Request-ID": "bench"})
            assert resp.status_code == 200, resp.text[:500]
            return _deserialize_response(resp, "pickle")

        warm, per, ok = [], [], True
        for i in range(args.warmup):
            t0 = time.perf_counter()
            out = call()
            warm.append(time.perf_counter() - t0)
            if i == 0:
                # the coordinator's own ranks come first, the remote pods' results follow in COMPLETION order
                # (remote_worker_pool.py:376-380 uses asyncio.as_completed): compare as a multiset of shards
                want = [c * 2 for c in x.chunk(n)]
                left = list(out)
                ok = len(out) == n
                for w in want:
                    hit = next((j for j, t in enumerate(left) if t.shape == w.shape and bool(torch.equal(t, w))), None)
                    ok = ok and hit is not None
                    if hit is not None:
                        left.pop(hit)
        for _ in range(args.steps):
            t0 = time.perf_counter()
            call()
            per.append(time.perf_counter() - t0)
        print("REFRESULT " + json.dumps({"shape": f"{n} pods x 1 rank (uvicorn on 127.0.0.k, HTTP between pods)",
                                         "per_call_s": per, "warm_s": warm, "ok": ok, "elems": args.elems, "ranks": n}),
              flush=True)
    finally:
        import signal

        for p in pods:      # each pod is its own session: take its spawned rank processes down with it
            try:
                os.killpg(p.pid, signal.SIGTERM)
            except (ProcessLookupError, PermissionError):
                pass
        for p in pods:
            try:
                p.wait(timeout=20)
            except subprocess.TimeoutExpired:
                try:
                    os.killpg(p.pid, signal.SIGKILL)
                except (ProcessLookupError, PermissionError):
                    pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", choices=["testclient", "pods"], default="testclient")
    ap.add_argument("--ranks", type=int, default=1)
    ap.add_argument("--elems", type=int, default=1 << 24)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    (run_testclient if args.shape == "testclient" else run_pods)(args)


if __name__ == "__main__":      # REQUIRED: the reference spawns its rank workers, which re-import __main__
    main()

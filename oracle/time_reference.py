"""Time the UNMODIFIED reference runtime beside the oracle port on the same machine (authoring container only).

TEST INFRASTRUCTURE.  The bench's CPU arm is the port (`OracleRuntime`), because /root/reference does not exist on the
GPU box.  This script shows the port is a FAIR stand-in: the real pod runtime (FastAPI TestClient → supervisor →
spawned ProcessWorkers, same recipe as make_golden.py) and the port run the same call here, wall-clock per call.
Usage:  python oracle/time_reference.py   → profiles/r1_reference_vs_port.json
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REFERENCE = "/root/reference/python_client"
N_ELEMS = 1 << 22          # 16 MiB fp32 arg, the bench's bounded sample
RANKS = 8


def child(out_path):
    os.environ["KT_LOG_STREAMING_ENABLED"] = "false"
    os.environ["KT_METRICS_ENABLED"] = "false"
    os.environ.update({
        "POD_NAMESPACE": "kubetorch", "POD_NAME": "timing-pod", "POD_IP": "localhost", "LOCAL_IPS": "localhost",
        "KT_SERVICE_NAME": "timing", "KT_FILE_PATH": REPO, "KT_MODULE_NAME": "oracle.cases",
        "KT_CLS_OR_FN_NAME": "double", "KT_INIT_ARGS": "null", "KT_ALLOWED_SERIALIZATION": "json,pickle",
        "KT_DISTRIBUTED_CONFIG": json.dumps({"distribution_type": "spmd", "num_proc": RANKS}),
    })
    import torch
    from fastapi.testclient import TestClient
    from kubetorch.resources.callables.utils import build_call_body
    from kubetorch.serving.http_server import app
    from kubetorch.serving.utils import _deserialize_response, _serialize_body

    x = torch.randn(N_ELEMS)
    small = torch.randn(256)
    res = {}
    with TestClient(app, raise_server_exceptions=False) as client:
        def call(t):
            body = _serialize_body(build_call_body(t), "pickle")
            resp = client.post("/double", json=body, headers={"X-Serialization": "pickle"})
            assert resp.status_code == 200
            return _deserialize_response(resp, "pickle")

        for name, t, n in (("16MiB", x, 3), ("1KiB", small, 50)):
            out = call(t)
            assert len(out) == RANKS and torch.equal(torch.cat(out), t * 2)
            t0 = time.perf_counter()
            for _ in range(n):
                call(t)
            res[name] = (time.perf_counter() - t0) / n
    with open(out_path, "w") as f:
        json.dump(res, f)


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--child":
        child(sys.argv[2])
        return
    import torch

    from oracle.make_golden import STUB
    from oracle.ref_dispatch import OracleRuntime

    work = tempfile.mkdtemp(prefix="kt_time_")
    with open(os.path.join(work, "websocket.py"), "w") as f:
        f.write(STUB)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([work, REFERENCE, REPO])
    env["HOME"] = work
    out_path = os.path.join(work, "ref.json")
    subprocess.run([sys.executable, os.path.abspath(__file__), "--child", out_path], env=env, check=True, cwd=work,
                   timeout=1200)
    with open(out_path) as f:
        ref = json.load(f)
    rt = OracleRuntime("oracle.cases", "double", RANKS)
    port = {}
    for name, t, n in (("16MiB", torch.randn(N_ELEMS), 3), ("1KiB", torch.randn(256), 50)):
        rt.call(t, serialization="pickle")
        t0 = time.perf_counter()
        for _ in range(n):
            rt.call(t, serialization="pickle")
        port[name] = (time.perf_counter() - t0) / n
    rt.close()
    nbytes = N_ELEMS * 4
    report = {
        "machine": f"authoring container, {os.cpu_count()} cores", "ranks": RANKS,
        "reference_runtime_unmodified": {"s_per_call_16MiB": ref["16MiB"], "GBps_arg_plus_result_16MiB": 2 * nbytes / ref["16MiB"] / 1e9,
                                         "calls_per_sec_1KiB": 1 / ref["1KiB"]},
        "oracle_port": {"s_per_call_16MiB": port["16MiB"], "GBps_arg_plus_result_16MiB": 2 * nbytes / port["16MiB"] / 1e9,
                        "calls_per_sec_1KiB": 1 / port["1KiB"]},
        "note": "the port omits the HTTP server, the 10 ms worker poll and the log/metric plumbing, so it is the FASTER of "
                "the two: speed-ups quoted against it are conservative",
    }
    out = os.path.join(REPO, "profiles", "r1_reference_vs_port.json")
    with open(out, "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()

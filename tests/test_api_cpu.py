"""not-gpu: the kt-compatible API on the CPU backends (in-process + rank processes) against the
reference's golden vectors and the recorded reference runtime behaviour."""
import asyncio
import json
import os
import threading
import time

import pytest
import torch

import kubetorch_b200 as kt
from conftest import REPO, resolve_args
from oracle import cases


def _base_name(exc):
    return next(c.__name__ for c in type(exc).__mro__ if c.__name__ != "RemoteException")


def _compute_for(cfg, allowed=None):
    comp = kt.Compute(cpus="1", allowed_serialization=(allowed.split(",") if allowed else None))
    if cfg["distribution_type"] != "local":
        # records taken on K real pods carry quorum_workers=K: the local backend emulates K nodes on 127.0.0.k
        comp.distribute(cfg["distribution_type"], workers=cfg.get("quorum_workers", 1), num_proc=cfg["num_proc"])
    return comp


@pytest.fixture(scope="module")
def deployed():
    """One deployment per (callable, config) for the whole module: rank processes start once."""
    cache = {}

    def get(callable_name, cfg, allowed=None):
        key = (callable_name, json.dumps(cfg, sort_keys=True), allowed)
        if key not in cache:
            obj = getattr(cases, callable_name)
            mod = kt.cls(obj, name=f"c-{len(cache)}") if isinstance(obj, type) else kt.fn(obj, name=f"c-{len(cache)}")
            cache[key] = mod.to(_compute_for(cfg, allowed))
        return cache[key]

    yield get
    for m in cache.values():
        m.teardown()


def _call(mod, rec_or_case, args):
    kwargs = dict(rec_or_case.get("kwargs") or {})
    kwargs["serialization"] = rec_or_case.get("serialization", "json")
    if rec_or_case.get("method"):
        return getattr(mod, rec_or_case["method"])(*args, **kwargs)
    return mod(*args, **kwargs)


def test_reference_asset_goldens_through_api(deployed):
    assets = json.load(open(os.path.join(REPO, "tests", "golden", "reference_assets.json")))
    for case in assets["cases"]:
        mod = deployed(case["callable"], case["distributed_config"])
        if "expected" in case:
            assert _call(mod, case, case["args"]) == case["expected"], case["name"]
        else:
            with pytest.raises(Exception) as ei:
                _call(mod, case, case["args"])
            assert _base_name(ei.value) == case["error"], case["name"]
            assert ei.value.status_code == case["error_code"], case["name"]
            assert ei.value.pod_name and "Traceback" in ei.value.remote_traceback


def _same(a, b):
    if isinstance(a, torch.Tensor):
        return isinstance(b, torch.Tensor) and a.dtype == b.dtype and a.shape == b.shape and \
            torch.equal(a.contiguous().reshape(-1).view(torch.uint8), b.contiguous().reshape(-1).view(torch.uint8))
    if isinstance(a, (list, tuple)):
        return isinstance(b, (list, tuple)) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, dict):
        return isinstance(b, dict) and a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    return a == b


def test_recorded_reference_runtime_through_api(golden, deployed):
    """Same requests as the reference runtime saw → same results / error type+message+status."""
    skip = {"number_count",            # depends on call history of one deployment
            "env_pytorch_4", "workers_any", "env_spmd_2"}  # POD_IPS/MASTER_ADDR are 127.0.0.x here, checked below
    n = 0
    for name, rec in golden["cases"].items():
        if name in skip or name.startswith(("mlp_", "number_state_", "torch_ddp_", "all_reduce_", "mp_all_reduce_")):
            continue  # the last three groups are replayed by test_recorded_process_groups_and_class_state
        mod = deployed(rec["callable"], rec["distributed_config"], rec["allowed"])
        args = resolve_args(golden, rec["args"])
        if rec["status_code"] == 200:
            assert _same(_call(mod, rec, args), rec["result"]), name
        else:
            with pytest.raises(Exception) as ei:
                _call(mod, rec, args)
            assert _base_name(ei.value) == rec["error"]["error_type"], name
            assert ei.value.args[0].split("\n\n")[0] == rec["error"]["message"], name
            assert ei.value.status_code == rec["status_code"], name
        n += 1
    assert n >= 25


def test_recorded_process_groups_and_class_state(golden):
    """Records that need their own deployment: real torch.distributed (gloo) ranks brought up from the env contract
    (each deployment owns MASTER_PORT 12345 while it lives), and an ORDERED call sequence on a stateful class."""
    def fresh(rec, name):
        obj = getattr(cases, rec["callable"])
        mod = kt.cls(obj, name=name) if isinstance(obj, type) else kt.fn(obj, name=name)
        cfg = rec["distributed_config"]
        comp = kt.Compute(cpus="1", allowed_serialization=rec["allowed"].split(","))
        # own rendezvous port: other deployments of this module may still hold the default 12345
        extra = {"port": 29541} if cfg["distribution_type"] == "pytorch" else {}
        return mod.to(comp.distribute(cfg["distribution_type"], workers=cfg.get("quorum_workers", 1),
                                      num_proc=cfg["num_proc"], **extra))

    # the last record was taken on TWO real pods x 2 ranks: one gloo group spanning both (here: two emulated nodes)
    for group in (("torch_ddp_valid_recorded", "torch_ddp_invalid_recorded"), ("all_reduce_rank_pt4",),
                  ("mp_all_reduce_rank_2x2",)):
        mod = fresh(golden["cases"][group[0]], f"pg-{group[0]}")
        try:
            for name in group:
                rec = golden["cases"][name]
                args = resolve_args(golden, rec["args"])
                if rec["status_code"] == 200:
                    assert _call(mod, rec, args) == rec["result"], name
                else:
                    with pytest.raises(Exception) as ei:
                        _call(mod, rec, args)
                    assert _base_name(ei.value) == rec["error"]["error_type"], name
                    assert ei.value.args[0].split("\n\n")[0] == rec["error"]["message"], name
                    assert ei.value.status_code == rec["status_code"], name
        finally:
            mod.teardown()
    seq = sorted(n for n in golden["cases"] if n.startswith("number_state_"))
    assert len(seq) == 4
    mod = fresh(golden["cases"][seq[0]], "number-state")
    try:
        for name in seq:   # count → add → add → count on the same per-rank instances
            rec = golden["cases"][name]
            assert _call(mod, rec, resolve_args(golden, rec["args"])) == rec["result"], name
    finally:
        mod.teardown()


def test_env_contract_matches_reference_modulo_addresses(golden, deployed):
    rec = golden["cases"]["env_pytorch_4"]
    got = deployed("env_report", rec["distributed_config"])()
    assert len(got) == 4
    for g, w in zip(got, rec["result"]):
        for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK", "NODE_RANK", "MASTER_PORT"):
            assert g[k] == w[k]
        assert g["MASTER_ADDR"] == "127.0.0.1" and g["POD_IPS"] == "127.0.0.1"
    spmd = deployed("env_report", golden["cases"]["env_spmd_2"]["distributed_config"])()
    assert [e["MASTER_ADDR"] for e in spmd] == [None, None] and [e["RANK"] for e in spmd] == ["0", "1"]


def test_multi_worker_world_and_selectors():
    f = kt.fn(cases.env_report, name="sel").to(kt.Compute(cpus=1).distribute("pytorch", workers=2, num_proc=2, port=29511))
    try:
        out = f()
        assert len(out) == 4  # workers × num_proc (tests/test_distributed.py:41)
        assert [e["RANK"] for e in out] == ["0", "1", "2", "3"]
        assert [e["NODE_RANK"] for e in out] == ["0", "0", "1", "1"] and out[0]["MASTER_PORT"] == "29511"
        assert out[0]["POD_IPS"] == "127.0.0.1,127.0.0.2"
        assert len(f(workers="any")) == 2 and [e["RANK"] for e in f(workers=[1])] == ["2", "3"]
        assert len(f(workers=["0", 1])) == 4 and len(f(workers=["127.0.0.2"])) == 2
        with pytest.raises(ValueError, match="Worker index 10 out of range. Valid range: 0-1"):
            f(workers=[10])
        with pytest.raises(ValueError, match="Worker IP '10.9.9.9' not found"):
            f(workers=["10.9.9.9"])
        assert len(f(restart_procs=True)) == 4
    finally:
        f.teardown()


def test_torch_distributed_launcher_gloo_all_reduce():
    """The launcher resolves rendezvous to local ranks; user code brings up the process group
    (tests/test_distributed.py:259-260: all_reduce of ranks at world 4 == 6.0)."""
    f = kt.fn(cases.all_reduce_rank, name="ar").to(
        kt.Compute(cpus=1).distribute("pytorch", workers=2, num_proc=2, port=29533))
    try:
        assert f() == [6.0, 6.0, 6.0, 6.0]
    finally:
        f.teardown()


def test_cls_state_persists_and_redeploy_resets():
    num = kt.cls(cases.Number, name="num")
    num.to(kt.Compute(cpus=1), init_args={"size": 7})
    assert num.add(1, 2) == 3 and num.add(2, 3) == 5 and num.count() == 2
    num.to(kt.Compute(cpus=1), init_args={"size": 7})  # redeploy → fresh instance
    assert num.count() == 0
    num.teardown()
    spmd = kt.cls(cases.Number, name="num-spmd").to(kt.Compute(cpus=1).distribute("spmd", num_proc=2))
    try:
        assert spmd.add(1, 2) == [3, 3] and spmd.count() == [1, 1]
    finally:
        spmd.teardown()


def test_reserved_kwargs_are_consumed_not_forwarded():
    f = kt.fn(cases.summer, name="rk").to(kt.Compute(cpus=1))
    assert f(1, 2, stream_logs=False, stream_metrics=False, serialization="json") == 3
    coro = f(1, 2, async_=True)
    assert asyncio.iscoroutine(coro) and asyncio.run(coro) == 3
    with pytest.raises(ValueError, match="Serialization must be"):
        f.serialization = "xml"
    f.teardown()
    with pytest.raises(ValueError, match="not deployed"):
        f(1, 2)


def test_json_mode_normalises_like_http_and_rejects_unserialisable():
    def pair():
        return (1, 2)

    pair.__module__ = cases.__name__
    setattr(cases, "pair", pair)
    f = kt.fn(cases.pair, name="pair").to(kt.Compute(cpus=1))
    assert f() == [1, 2]  # a tuple crosses JSON as a list, as through the reference's HTTP hop
    g = kt.fn(cases.spmd_identity, name="ident-json").to(kt.Compute(cpus=1))
    with pytest.raises(TypeError):
        g(torch.ones(2))  # tensors need serialization="pickle"
    assert torch.equal(g(torch.ones(2), serialization="pickle"), torch.ones(2))
    f.teardown(); g.teardown()


def test_serialization_formats_with_pydantic_models_like_the_reference_suite():
    """kt tests/test_deployment_fixtures.py:116-205: ints and Pydantic models through pickle, class methods with model
    lists, the module-level `.serialization` setting, per-call overrides, and serialization="none"."""
    from pydantic import BaseModel

    comp = kt.Compute(cpus=".01", allowed_serialization=["json", "pickle", "none"])
    remote_fn = kt.fn(cases.model_summer, name="ser-fn").to(comp)
    remote_cls = kt.cls(cases.OSInfo, name="ser-cls").to(
        kt.Compute(cpus=".01", allowed_serialization=["json", "pickle", "none"]), init_args={"size": 3})
    try:
        assert remote_fn(2, 3, serialization="pickle") == 5
        out = remote_fn(cases.PairModel(name="test_a", value=42), cases.PairModel(name="test_b", value=42),
                        serialization="pickle")
        assert isinstance(out, cases.PairModel) and out.name == "sum_result" and out.value == 84
        cpu_count = remote_cls.cpu_count(serialization="pickle")
        assert isinstance(cpu_count, int)
        reqs = [cases.OSInfoRequest(method=m) for m in ("uname", "cpu_count", "getpid")]
        res = remote_cls.os_info(reqs, serialization="pickle")
        assert isinstance(res, list) and all(isinstance(r, BaseModel) for r in res)
        assert [r.name for r in res] == ["uname", "cpu_count", "getpid"] and res[1].value == str(cpu_count)
        with pytest.raises(TypeError):  # models are not JSON-serialisable: the json route refuses them like httpx would
            remote_fn(cases.PairModel(name="a", value=1), cases.PairModel(name="b", value=2), serialization="json")
        # module-level default, still overridable per call
        original = remote_fn.serialization
        remote_fn.serialization = "pickle"
        try:
            out = remote_fn(cases.PairModel(name="test_a", value=42), cases.PairModel(name="test_b", value=42))
            assert isinstance(out, cases.PairModel) and out.value == 84
            assert remote_fn(3, 4, serialization="json") == 7
            assert remote_fn(3, 4, serialization="none") == 7
        finally:
            remote_fn.serialization = original
        # the same models across real rank processes (pickled through the worker pipes), one result per rank
        spmd = kt.fn(cases.model_summer, name="ser-fn-spmd").to(
            kt.Compute(cpus="1", allowed_serialization=["json", "pickle"]).distribute("spmd", workers=1, num_proc=2))
        try:
            outs = spmd(cases.PairModel(name="x", value=40), cases.PairModel(name="y", value=2), serialization="pickle")
            assert [type(o) for o in outs] == [cases.PairModel] * 2 and [o.value for o in outs] == [42, 42]
        finally:
            spmd.teardown()
        remote_cls.serialization = "pickle"
        try:
            res = remote_cls.os_info([cases.OSInfoRequest(method="cpu_count")])
            assert isinstance(res[0], BaseModel) and res[0].value == str(cpu_count)
            assert isinstance(remote_cls.size_minus_cpus(serialization="json"), int)
            assert isinstance(remote_cls.size_minus_cpus(serialization="none"), int)
        finally:
            remote_cls.serialization = "json"
    finally:
        remote_fn.teardown()
        remote_cls.teardown()


def test_async_callables_overlap_and_sync_calls_run_concurrently():
    f = kt.fn(cases.async_summer, name="as").to(kt.Compute(cpus=1).distribute("spmd", num_proc=1))
    try:
        t0 = time.time()
        outs = []
        ths = [threading.Thread(target=lambda: outs.append(f(1, 2, sleep_time=0.4))) for _ in range(4)]
        [t.start() for t in ths]; [t.join() for t in ths]
        assert outs == [[3]] * 4 and time.time() - t0 < 1.2  # 4 × 0.4 s overlapped on one event loop
    finally:
        f.teardown()


def test_worker_death_maps_to_pod_terminated_error():
    def die():
        os._exit(3)

    die.__module__ = cases.__name__
    setattr(cases, "die", die)
    # the spawned ranks import oracle.cases fresh, so use a callable that exists there
    f = kt.fn(cases.raise_value_error, name="dead").to(kt.Compute(cpus=1).distribute("spmd", num_proc=2))
    try:
        f._supervisor.pool._procs[1].terminate()
        f._supervisor.pool._procs[1].join()
        with pytest.raises(kt.PodTerminatedError):
            f("x")
    finally:
        f.teardown()


def test_compute_and_distribute_config_surface():
    c = kt.Compute(gpus=8, memory="64Gi", image=None, launch_timeout=300)
    assert c.distributed_config == {}
    c.distribute("pytorch", workers=4, num_proc=8, port=1234)
    assert c.distributed_config == {"distribution_type": "pytorch", "quorum_timeout": 300, "quorum_workers": 4,
                                    "num_proc": 8, "port": 1234}
    assert c.replicas == 4
    with pytest.raises(ValueError, match="Workers must be an integer"):
        kt.Compute(cpus=1).distribute("spmd", workers=[1, 2])
    with pytest.raises(ValueError, match="non-serializable"):
        kt.Compute(cpus=1).distribute("spmd", workers=1, bad=object())
    with pytest.raises(ValueError, match="Unsupported distribution type"):
        kt.fn(cases.summer, name="bad").to(kt.Compute(cpus=1).distribute("mpi"))
    for name in ("PodTerminatedError", "WorkerMembershipChanged", "StartupError", "ImagePullError"):
        assert name in kt.EXCEPTION_REGISTRY and kt.EXCEPTION_REGISTRY[name].__module__ == "kubetorch_b200"


def test_cluster_side_resources_are_accepted_at_call_sites():
    """Programs written against the reference construct images / secrets / volumes next to their compute; on the local
    route they are inert, except Image.set_env_vars which reaches the rank processes."""
    img = kt.images.Debian().pip_install(["pytest", "fastapi"]).set_env_vars({"KTB_TEST_FLAG": "from-image"})
    assert isinstance(img, kt.Image) and img.steps == [("pip_install", ["pytest", "fastapi"])]
    assert kt.images.Python312().image_id == "python:3.12-slim" and kt.images.pytorch().name == "pytorch2312py3"
    comp = kt.Compute(cpus=".01", gpu_anti_affinity=True, launch_timeout=300, image=img, secrets=[kt.secret(name="hf")],
                      volumes=[kt.Volume(name="data", size="1Gi", mount_path="/data")], env_vars={"OTHER": "1"})
    assert comp.env_vars == {"KTB_TEST_FLAG": "from-image", "OTHER": "1"}
    with pytest.raises(ValueError, match="Either name or provider"):
        kt.secret()
    with pytest.raises(NotImplementedError):
        comp.autoscale(min_scale=1)

    f = kt.fn(cases.env_get, name="img-env").to(comp.distribute("spmd", workers=1, num_proc=2))
    try:
        assert f("KTB_TEST_FLAG") == ["from-image"] * 2 and f("OTHER") == ["1", "1"]
    finally:
        f.teardown()


def test_config_precedence_setter_env_file_default(tmp_path, monkeypatch):
    """kt/config.py:13-25,76-95: explicit setter > KT_* env > ~/.kt/config.yaml > default."""
    from kubetorch_b200.config import KubetorchConfig

    cfg_file = tmp_path / "config.yaml"
    cfg_file.write_text("namespace: from-file\nstream_logs: false\n")
    monkeypatch.setattr(KubetorchConfig, "CONFIG_FILE", str(cfg_file))
    monkeypatch.delenv("KT_NAMESPACE", raising=False)
    monkeypatch.delenv("KT_STREAM_LOGS", raising=False)
    cfg = KubetorchConfig()
    assert cfg.namespace == "from-file" and cfg.stream_logs is False and cfg.stream_metrics is False
    monkeypatch.setenv("KT_NAMESPACE", "from-env")
    assert cfg.namespace == "from-env"
    cfg.namespace = "explicit"
    assert cfg.namespace == "explicit"
    with pytest.raises(AttributeError):
        cfg.no_such_setting


def test_tensor_wire_split_and_join_roundtrip():
    import pickle

    from kubetorch_b200.serving.tensor_wire import TensorRef, collect_refs, join_tensors, split_tensors

    a, b = torch.arange(6.0).reshape(2, 3), torch.ones(4, dtype=torch.int64)
    payload = ([a, 3, "s"], {"k": (b, {"deep": a}), "n": None})
    leaves = []
    skel = split_tensors(payload, leaves, lambda t: True)
    assert len(leaves) == 3 and leaves[0] is a and leaves[1] is b
    skel2 = pickle.loads(pickle.dumps(skel))  # the header is what crosses the pipe
    refs = []
    collect_refs(skel2, refs)
    assert [(r.index, r.dtype, r.shape) for r in refs] == [(0, "float32", (2, 3)), (1, "int64", (4,)), (2, "float32", (2, 3))]
    back = join_tensors(skel2, leaves)
    assert back[0][1:] == [3, "s"] and back[0][0] is a and back[1]["k"][0] is b and back[1]["k"][1]["deep"] is a
    assert isinstance(skel[0][0], TensorRef) and split_tensors(5, [], lambda t: True) == 5


def test_gpu_compute_without_gpu_falls_back_to_plain_rank_processes_for_python_callables():
    """kt.Compute(gpus=N).distribute("spmd") with an arbitrary callable is the process route; on a box without
    CUDA it still runs (tensors travel pickled) — only @mapped/b200 callables require the device library."""
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    f = kt.fn(cases.double, name="gpu-less").to(
        kt.Compute(gpus=2, allowed_serialization=["json", "pickle"]).distribute("spmd", workers=1, num_proc=2))
    try:
        x = torch.arange(7, dtype=torch.float32)
        out = f(x, serialization="pickle")
        assert torch.equal(torch.cat(out), x * 2)
    finally:
        f.teardown()


def test_data_store_surface_without_gpu():
    assert kt.BroadcastWindow(world_size=4, pack=True).to_dict()["pack"] is True
    with pytest.raises(ValueError, match="at least one of"):
        kt.BroadcastWindow()
    with pytest.raises(ValueError, match="src is required"):
        kt.put("k")
    with pytest.raises(NotImplementedError):
        kt.put("k", src="./some/dir")          # filesystem keys belong to the Kubernetes rsync store
    with pytest.raises(NotImplementedError):
        kt.get("k", dest="./some/dir")
    with pytest.raises(kt.DataStoreError):
        kt.rm("never-put")
    assert kt.ls("never-put") == []


def test_endpoint_must_name_the_deployed_callable():
    from kubetorch_b200.serving.supervisors import Request

    f = kt.fn(cases.summer, name="guard").to(kt.Compute(cpus=1))
    try:
        with pytest.raises(Exception, match="Callable 'other' not found in metadata configuration. Found 'summer' instead"):
            f._supervisor.call(Request({"X-Serialization": "json"}), "other", None, {"args": [1, 2], "kwargs": {}})
        with pytest.raises(Exception) as ei:
            f._client().call_method("local://guard/other", body={"args": [1, 2], "kwargs": {}})
        assert ei.value.status_code == 404
    finally:
        f.teardown()


def test_fastpickle_roundtrips_and_sends_only_addressed_bytes():
    """Pipes between coordinator and ranks: plain CPU tensors travel as (dtype, shape, raw bytes); everything else
    takes the stock reduction.  A view must not drag its whole storage along."""
    import pickle

    from kubetorch_b200.serving import fastpickle as F

    big = torch.randn(1 << 18)
    objs = [torch.randn(256), torch.randn(7, 3).bfloat16(), torch.tensor(3.5), torch.empty(0, 4), big[5:261],
            torch.randn(4, 5).t(), torch.randint(0, 2, (9,)).bool(), torch.randn(5).half(),
            torch.randn(3, requires_grad=True), torch.nn.Parameter(torch.randn(2)),
            {"a": [torch.ones(2), ("x", torch.zeros(1, dtype=torch.int64))], "b": 5, "c": None}]

    def same(a, b):
        if isinstance(a, torch.Tensor):
            return type(a) is type(b) and a.dtype == b.dtype and a.shape == b.shape and \
                a.requires_grad == b.requires_grad and torch.equal(a, b)
        if isinstance(a, dict):
            return a.keys() == b.keys() and all(same(a[k], b[k]) for k in a)
        if isinstance(a, (list, tuple)):
            return type(a) is type(b) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        return a == b

    for o in objs:
        assert same(o, F.loads(F.dumps(o)))
    assert len(F.dumps(big[5:261])) < 2048 < len(pickle.dumps(big[5:261]))
    got = F.loads(F.dumps(torch.zeros(4)))
    got += 1  # results are writable tensors
    assert got.tolist() == [1.0] * 4


def test_concurrent_large_payloads_do_not_interleave_on_the_rank_pipes():
    """4 caller threads x 1 MiB+ tensor payloads on the same rank pipes (Connection.send_bytes writes header and
    payload separately above 16 KiB, so unsynchronised writers corrupt the stream)."""
    remote = kt.fn(cases.spmd_identity, name="c-concurrent").to(
        kt.Compute(cpus="1", allowed_serialization=["json", "pickle"]).distribute("spmd", workers=1, num_proc=2))
    try:
        xs = [torch.full((300_000 + 1000 * i,), float(i)) for i in range(4)]
        out = [None] * 4
        errs = []

        def work(i):
            try:
                for _ in range(5):
                    out[i] = remote(xs[i], serialization="pickle")
            except BaseException as e:  # noqa: BLE001
                errs.append(e)

        ths = [threading.Thread(target=work, args=(i,)) for i in range(4)]
        [t.start() for t in ths]
        [t.join(timeout=120) for t in ths]
        assert not errs, errs
        for i in range(4):
            assert len(out[i]) == 2 and all(torch.equal(o, xs[i]) for o in out[i])
    finally:
        remote.teardown()


def test_partial_deployments_spread_over_numa_nodes():
    """B200Supervisor._pick_devices: N of the visible GPUs, GPU 0 first, spread over the sockets when N < visible."""
    from kubetorch_b200.serving.b200_supervisor import B200Supervisor

    class Box:
        def __init__(self, nodes):
            self.nodes = nodes

        def device_count(self):
            return len(self.nodes)

        def device_numa_node(self, d):
            return self.nodes[d]

    sup = B200Supervisor()
    sup.ops = Box([0, 0, 0, 0, 1, 1, 1, 1])
    assert [sup._pick_devices(n) for n in (1, 2, 3, 4, 6, 8)] == \
        [[0], [0, 4], [0, 1, 4], [0, 1, 4, 5], [0, 1, 2, 4, 5, 6], list(range(8))]
    sup.ops = Box([1, 1, 1, 1, 0, 0, 0, 0])                # node numbering is arbitrary: GPU 0's socket comes first
    assert sup._pick_devices(4) == [0, 1, 4, 5]
    sup.ops = Box([0, 0, 0, 0])                            # one socket: first N
    assert sup._pick_devices(2) == [0, 1]
    sup.ops = Box([-1] * 8)                                # no topology information
    assert sup._pick_devices(4) == [0, 1, 2, 3]


def test_map_falls_back_to_a_loop_of_calls_on_cpu_backends():
    remote = kt.fn(cases.summer, name="c-map").to(kt.Compute(cpus="1"))
    try:
        assert remote.map([]) == []
        with pytest.raises(TypeError):
            remote.map([(1, 2)])                           # one positional argument per item, like remote(x)
    finally:
        remote.teardown()
    ident = kt.fn(cases.spmd_identity, name="c-map2").to(
        kt.Compute(cpus="1", allowed_serialization=["json", "pickle"]).distribute("spmd", workers=1, num_proc=2))
    try:
        xs = [torch.arange(5.0) + i for i in range(3)]
        out = ident.map(xs, serialization="pickle")
        assert len(out) == 3 and all(len(o) == 2 and torch.equal(o[0], x) and torch.equal(o[1], x) for o, x in zip(out, xs))
    finally:
        ident.teardown()


def test_broadcast_window_rendezvous_times_out_cleanly_without_a_putter(tmp_path, monkeypatch):
    """The file rendezvous of a cross-process BroadcastWindow (no GPU needed up to the pull): a window that cannot
    close raises on the participant and leaves no join file behind."""
    from kubetorch_b200 import data_store

    monkeypatch.setenv("KTB_STORE_DIR", str(tmp_path))
    bw = kt.BroadcastWindow(world_size=3, timeout=0.2, group_id="cpu-bw")
    with pytest.raises(kt.DataStoreError, match=r"timed out with 0 putter\(s\) and 1 getter\(s\)"):
        data_store._join_shared("k", [("", torch.zeros(1))], bw, "get")
    gdir = [p for p in tmp_path.iterdir() if p.name.startswith("bw_")]
    assert gdir and not any(f.name.endswith(".join") for f in gdir[0].iterdir())

import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
# spawned rank processes import the user modules (oracle.cases, tests.*) by name
os.environ["PYTHONPATH"] = os.pathsep.join([REPO] + [p for p in os.environ.get("PYTHONPATH", "").split(os.pathsep) if p])


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import torch

    from oracle.golden_inputs import make_inputs, tensor_sha256

    fx = torch.load(os.path.join(REPO, "tests", "golden", "ref_runtime.pt"), weights_only=False)
    inputs = make_inputs()
    for k, v in inputs.items():  # regenerated inputs must be the tensors the reference saw
        assert tensor_sha256(v) == fx["input_sha256"][k], f"golden input {k} does not regenerate bit-exactly"
    fx["all_inputs"] = inputs
    return fx


def resolve_args(fx, arg_spec):
    """Replace "@name" strings (at any depth of lists/dicts) by the named golden input tensor."""
    if isinstance(arg_spec, str) and arg_spec.startswith("@"):
        return fx["all_inputs"][arg_spec[1:]]
    if isinstance(arg_spec, list):
        return [resolve_args(fx, v) for v in arg_spec]
    if isinstance(arg_spec, dict):
        return {k: resolve_args(fx, v) for k, v in arg_spec.items()}
    return arg_spec


def mapped_copy(fn, *a, **k):
    """@kt.mapped applied to a COPY of fn, so shared oracle callables are never mutated by a test."""
    import functools
    import types

    import kubetorch_b200 as kt

    clone = types.FunctionType(fn.__code__, fn.__globals__, fn.__name__, fn.__defaults__, fn.__closure__)
    clone = functools.update_wrapper(clone, fn)
    clone.__dict__.pop("__ktb_mapped__", None)
    del clone.__wrapped__
    return kt.mapped(*a, **k)(clone)

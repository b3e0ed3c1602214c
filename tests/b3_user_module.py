"""User module deployed behind the supervisor seam in tests/test_b3_seam.py (REFERENCE server, CPU stub) and
tests/test_gpu_api.py::test_b3_* (real kernels): the parity callables of oracle/cases.py — plain kubetorch-style
SPMD functions that shard by RANK/WORLD_SIZE — plus the ONE new line per function that declares its device op.
Loaded the way the reference loads user code: by name, from KT_FILE_PATH / KT_MODULE_NAME / KT_CLS_OR_FN_NAME."""
import functools
import types

import kubetorch_b200 as ktb
from oracle import cases


def _mapped(fn, *a, **k):
    """@ktb.mapped on a COPY of the oracle callable (the shared function object stays undecorated)."""
    clone = types.FunctionType(fn.__code__, fn.__globals__, fn.__name__, fn.__defaults__, fn.__closure__)
    clone = functools.update_wrapper(clone, fn)
    clone.__dict__.pop("__ktb_mapped__", None)
    del clone.__wrapped__
    return ktb.mapped(*a, **k)(clone)


double = _mapped(cases.double, "scale", alpha=2.0)
identity = _mapped(cases.identity, "identity")
scale = _mapped(cases.scale, "scale", alpha="alpha")
affine = _mapped(cases.affine, "affine", alpha="alpha", beta="beta")
shard_sum = _mapped(cases.shard_sum, "affine", alpha="alpha", beta="beta", reduce="sum")


@ktb.mapped("scale", alpha=2.0)
def not_really_double(x):
    """A WRONG declaration (the body triples): the deploy-time self-check must refuse it."""
    return cases._shard(x) * 3


class Scaler:
    """kt.cls with a mapped method: one instance per deployment, constructed from KT_INIT_ARGS."""

    def __init__(self, tag="none"):
        self.tag = tag

    @ktb.mapped("scale", alpha=3.0)
    def triple(self, x):
        return cases._shard(x) * 3

"""@kt.compute / @kt.distribute / @kt.async_ (kt/resources/compute/decorators.py:31-138): build the
Fn/Cls for the decorated object with its Compute attached; `.deploy()` (or first `.to`) binds it."""
from __future__ import annotations

import inspect

from .compute import Compute


def _as_module(obj):
    from .callables import Module, cls, fn

    if isinstance(obj, Module):
        return obj
    return cls(obj) if inspect.isclass(obj) else fn(obj)


def compute(get_if_exists: bool = False, reload_prefixes=None, **kwargs):
    def deco(obj):
        mod = _as_module(obj)
        mod.compute = Compute(**kwargs)
        mod.get_if_exists = get_if_exists
        mod.reload_prefixes = reload_prefixes or []
        return mod

    return deco


def distribute(distribution_type: str = None, workers: int = None, **kwargs):
    def deco(obj):
        mod = _as_module(obj)
        if mod.compute is None:
            raise ValueError("@kt.distribute must be applied above @kt.compute")
        mod.compute.distribute(distribution_type, workers=workers, **kwargs)
        return mod

    return deco


def async_(obj):
    mod = _as_module(obj)
    mod.async_ = True
    return mod


def autoscale(**kwargs):
    """@kt.autoscale (kt/resources/compute/decorators.py): Knative autoscaling has no meaning on a fixed set of local
    GPUs; the decorator raises the same error Compute.autoscale() does, at decoration time."""
    def deco(obj):
        mod = _as_module(obj)
        if mod.compute is None:
            raise ValueError("@kt.autoscale must be applied above @kt.compute")
        mod.compute.autoscale(**kwargs)
        return mod

    return deco

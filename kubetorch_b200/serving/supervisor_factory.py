"""supervisor_factory(distribution_type, **config) — same seam and names as the reference
(kt/serving/supervisor_factory.py:11-58), plus the new "b200" type."""
from __future__ import annotations

from .b200_supervisor import B200Supervisor
from .supervisors import ExecutionSupervisor, SPMDSupervisor

SPMD_TYPES = ("spmd", "pytorch", "jax", "tensorflow", "tf")


def supervisor_factory(distribution_type, *args, **kwargs):
    if distribution_type == "local":
        return ExecutionSupervisor(*args, **kwargs)
    if distribution_type == "b200":
        return B200Supervisor(*args, **kwargs)
    if distribution_type in ("ray", "monarch"):
        raise ValueError(
            f"distribution type '{distribution_type}' launches a third-party runtime on Kubernetes pods and is "
            "outside the local-B200 route"
        )
    if distribution_type is None or distribution_type in SPMD_TYPES:
        dt = "tensorflow" if distribution_type == "tf" else (distribution_type or "spmd")
        if kwargs.pop("gpu_arenas", False):
            from .gpu_spmd import GpuSPMDSupervisor

            return GpuSPMDSupervisor(distribution_type=dt, *args, **kwargs)
        kwargs.pop("devices", None)
        return SPMDSupervisor(distribution_type=dt, *args, **kwargs)
    raise ValueError(f"Unsupported distribution type: {distribution_type}")

"""kubetorch_b200 — a Blackwell-native dispatch backend for kubetorch's data-parallel remote-call
path.  `import kubetorch_b200 as kt` is a drop-in for the subset of `import kubetorch as kt` on
that path: kt.Compute / kt.fn / kt.cls / .to() / .distribute() / remote __call__
(exports mirror kt/__init__.py:1-36), with `kt.Compute(gpus=N)` bound to N local B200s.
"""
from . import distributed  # noqa: F401
from .data_store import BroadcastWindow, get, ls, put, rm  # noqa: F401
from .config import DebugConfig, LoggingConfig, MetricsConfig, config  # noqa: F401
from .exceptions import (  # noqa: F401
    EXCEPTION_REGISTRY,
    ControllerRequestError,
    DataStoreError,
    ImagePullError,
    KnativeServiceConflictError,
    KubernetesCredentialsError,
    NotebookError,
    PodContainerError,
    PodTerminatedError,
    ResourceNotAvailableError,
    RsyncError,
    SecretNotFound,
    SerializationError,
    ServiceHealthError,
    ServiceTimeoutError,
    StartupError,
    VersionMismatchError,
    WorkerMembershipChanged,
)
from .mapped import mapped, mapped_spec  # noqa: F401
from .resources.callables import Cls, Fn, Module, cls, fn  # noqa: F401
from .resources.compute import Compute  # noqa: F401
from .resources.decorators import async_, autoscale, compute, distribute  # noqa: F401
from .resources.inert import Image, Secret, Volume, images, secret  # noqa: F401  (call-site stand-ins, see inert.py)

for _exc in EXCEPTION_REGISTRY.values():
    _exc.__module__ = "kubetorch_b200"

__version__ = "0.1.0"

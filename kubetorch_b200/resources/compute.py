"""kt.Compute — constructor signature and .distribute() of the reference
(kt/resources/compute/compute.py:34-69, 1569-1604, 2596-2694), bound to LOCAL resources:
`Compute(cpus=…)` → host processes, `Compute(gpus=N)` → N local B200s.  Kubernetes-only arguments
(image, secrets, volumes, tolerations, …) are accepted and kept as inert attributes so existing
call sites run unchanged.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Union

from ..config import LoggingConfig

DEFAULT_ALLOWED_SERIALIZATION = ["json", "pickle"]


class Compute:
    def __init__(
        self,
        cpus: Union[str, int] = None,
        memory: str = None,
        disk_size: str = None,
        gpus: Union[str, int] = None,
        gpu_type: str = None,
        priority_class_name: str = None,
        gpu_memory: str = None,
        namespace: str = None,
        image=None,
        labels: Dict = None,
        annotations: Dict = None,
        volumes: List = None,
        node_selector: Dict = None,
        service_template: Dict = None,
        tolerations: List[Dict] = None,
        env_vars: Dict = None,
        secrets: List = None,
        freeze: bool = False,
        kubeconfig_path: str = None,
        service_account_name: str = None,
        image_pull_policy: str = None,
        inactivity_ttl: str = None,
        gpu_anti_affinity: bool = None,
        launch_timeout: int = None,
        working_dir: str = None,
        shared_memory_limit: str = None,
        allowed_serialization: Optional[List[str]] = None,
        replicas: int = None,
        logging_config: LoggingConfig = None,
        queue_name: str = None,
        selector: Dict[str, str] = None,
        endpoint=None,
        _skip_template_init: bool = False,
    ):
        self.cpus = cpus
        self.memory = memory
        self.disk_size = disk_size
        self.gpus = int(gpus) if gpus not in (None, "") else None
        self.gpu_type = gpu_type
        self.gpu_memory = gpu_memory
        self.namespace = namespace or "default"
        self.image = image
        self.labels = labels or {}
        self.annotations = annotations or {}
        self.volumes = volumes or []
        self.node_selector = node_selector
        self.service_template = service_template
        self.tolerations = tolerations
        self.env_vars = dict(env_vars or {})
        if image is not None and getattr(image, "env_vars", None):
            # Image.set_env_vars is the one image setting with a local meaning: exported to the rank processes
            self.env_vars = {**image.env_vars, **self.env_vars}
        self.secrets = secrets or []
        self.freeze = freeze
        self.kubeconfig_path = kubeconfig_path
        self.service_account_name = service_account_name
        self.image_pull_policy = image_pull_policy
        self.inactivity_ttl = inactivity_ttl
        self.gpu_anti_affinity = gpu_anti_affinity
        self.launch_timeout = launch_timeout or 900
        self.working_dir = working_dir
        self.shared_memory_limit = shared_memory_limit
        self.priority_class_name = priority_class_name
        self.allowed_serialization = list(allowed_serialization) if allowed_serialization else None
        self.replicas = replicas or 1
        self.logging_config = logging_config or LoggingConfig()
        self.queue_name = queue_name
        self.selector = selector
        self.endpoint = endpoint
        self.service_name = None
        self._distributed_config = None
        self._autoscaling_config = None

    # ---- serialization allow-list (http_server.py:1777-1782 via KT_ALLOWED_SERIALIZATION) ---------------
    @property
    def allowed_serialization_str(self) -> str:
        return ",".join(self.allowed_serialization or DEFAULT_ALLOWED_SERIALIZATION)

    # ---- distributed config ---------------------------------------------------------------------------
    @property
    def distributed_config(self) -> dict:
        return self._distributed_config or {}

    @distributed_config.setter
    def distributed_config(self, config: dict):
        workers = config.get("workers")
        config["distribution_type"] = config.get("distribution_type", "spmd")
        config["quorum_timeout"] = config.get("quorum_timeout", self.launch_timeout)
        config["quorum_workers"] = config.get("quorum_workers", workers or self.replicas)
        self.replicas = workers or config["quorum_workers"]
        bad = []
        for key, value in config.items():
            try:
                json.dumps(value)
            except (TypeError, ValueError) as e:
                bad.append(f"'{key}': {type(value).__name__} - {e}")
        if bad:
            raise ValueError(
                f"Distributed config contains non-serializable values: {', '.join(bad)}. "
                f"All values must be JSON serializable (strings, numbers, booleans, lists, dicts)."
            )
        self._distributed_config = config

    @property
    def autoscaling_config(self):
        return self._autoscaling_config

    def distribute(self, distribution_type: str = None, workers: int = None, quorum_timeout: int = None,
                   quorum_workers: int = None, monitor_members: bool = None, **kwargs):
        """Configure the ranks of each call: `workers` (emulated pods) × `num_proc` ranks per worker.

        distribution_type: "spmd" (default), "pytorch", "jax", "tensorflow" — rank processes with the
        reference's env contract — or "b200": the device-kernel backend for @kt.mapped callables.
        """
        if self.autoscaling_config:
            raise ValueError(
                "Cannot use both .distribute() and .autoscale() on the same compute instance. "
                "Use .distribute() for fixed replicas with distributed training, or .autoscale() for auto-scaling services."
            )
        quorum_workers = quorum_workers or workers
        cfg = {
            "distribution_type": distribution_type or "spmd",
            "quorum_timeout": quorum_timeout or self.launch_timeout,
            "quorum_workers": quorum_workers,
        }
        if monitor_members is not None:
            cfg["monitor_members"] = monitor_members
        cfg.update(kwargs)
        if workers:
            if not isinstance(workers, int):
                raise ValueError("Workers must be an integer. List of <integer, Compute> pairs is not yet supported")
            self.replicas = workers
        self.distributed_config = cfg
        return self

    def autoscale(self, **kwargs):
        raise NotImplementedError("autoscaling is a Kubernetes (Knative) feature; the local-B200 route has fixed ranks")

    def __repr__(self):
        return (f"Compute(cpus={self.cpus!r}, gpus={self.gpus!r}, replicas={self.replicas}, "
                f"distributed_config={self.distributed_config})")

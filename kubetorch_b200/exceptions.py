"""Exception types user code scripts against, with the reference's names, attributes and wire state.

Mirrors kt/serving/utils.py:107-263 (StartupError, PodTerminatedError, WorkerMembershipChanged) and
the registry of kt/__init__.py:43-60 used to rebuild typed exceptions on the caller side
(kt/serving/http_client.py:87-175).  K8s-only error types are kept as plain Exception subclasses so
`except kt.ImagePullError` in user code still imports.
"""
from __future__ import annotations

from typing import List


class StartupError(Exception):
    pass


class SerializationError(Exception):
    pass


class PodTerminatedError(Exception):
    """On the local-B200 route a "pod" is a rank process; its death (signal, CUDA fault) maps here."""

    def __init__(self, pod_name: str = "unknown", reason: str = "Unknown", status_code: int = 503,
                 events: List[dict] = None):
        self.pod_name = pod_name
        self.reason = reason
        self.status_code = status_code
        self.events = events or []
        super().__init__(str(self))

    def __getstate__(self):
        events = []
        for ev in self.events:
            ev = dict(ev)
            ts = ev.get("timestamp")
            if hasattr(ts, "isoformat"):
                ev["timestamp"] = ts.isoformat()
            events.append(ev)
        return {"pod_name": self.pod_name, "reason": self.reason, "status_code": self.status_code, "events": events}

    def __setstate__(self, state):
        self.pod_name, self.reason = state["pod_name"], state["reason"]
        self.status_code, self.events = state["status_code"], state["events"]

    @classmethod
    def from_dict(cls, state):
        return cls(pod_name=state.get("pod_name", "unknown"), reason=state.get("reason", "Unknown"),
                   status_code=state.get("status_code", 503), events=state.get("events", []))

    @property
    def evicted(self) -> bool:
        return self.reason == "Evicted" or any("Evicted" in e["reason"] for e in self.events)

    @property
    def oom_killed(self) -> bool:
        return self.reason == "OOMKilled" or any("OOMKilled" in e["reason"] for e in self.events)

    def __str__(self):
        text = f"\nPod Name: {self.pod_name}\nReason: {self.reason}\nStatus Code: {self.status_code}\n"
        if self.events:
            text += "Recent Events:\n" + "\n".join(f"{e['timestamp']} {e['reason']}: {e['message']}" for e in self.events)
        return text


class WorkerMembershipChanged(Exception):
    """Membership is static on one box; kept for API parity (never raised by the local backends)."""

    def __init__(self, added_ips: set = None, removed_ips: set = None, previous_ips: set = None,
                 current_ips: set = None, message: str = None):
        self.added_ips = set(added_ips or ())
        self.removed_ips = set(removed_ips or ())
        self.previous_ips = set(previous_ips or ())
        self.current_ips = set(current_ips or ())
        if message is None:
            if self.removed_ips:
                message = f"Critical: {len(self.removed_ips)} worker(s) removed during execution: {self.removed_ips}"
            elif self.added_ips:
                message = f"Warning: {len(self.added_ips)} worker(s) added during execution: {self.added_ips}"
            else:
                message = "Worker membership changed"
        super().__init__(message)

    @property
    def is_critical(self) -> bool:
        return bool(self.removed_ips)

    def __getstate__(self):
        return {"message": str(self), "added_ips": list(self.added_ips), "removed_ips": list(self.removed_ips),
                "previous_ips": list(self.previous_ips), "current_ips": list(self.current_ips)}

    @classmethod
    def from_dict(cls, data):
        return cls(added_ips=set(data.get("added_ips", [])), removed_ips=set(data.get("removed_ips", [])),
                   previous_ips=set(data.get("previous_ips", [])), current_ips=set(data.get("current_ips", [])))


def _k8s_only(name: str):
    return type(name, (Exception,), {"__doc__": f"{name}: Kubernetes-route error type, importable for drop-in code."})


ControllerRequestError = _k8s_only("ControllerRequestError")
ImagePullError = _k8s_only("ImagePullError")
KubernetesCredentialsError = _k8s_only("KubernetesCredentialsError")
KnativeServiceConflictError = _k8s_only("KnativeServiceConflictError")
PodContainerError = _k8s_only("PodContainerError")
ResourceNotAvailableError = _k8s_only("ResourceNotAvailableError")
RsyncError = _k8s_only("RsyncError")
SecretNotFound = _k8s_only("SecretNotFound")
ServiceHealthError = _k8s_only("ServiceHealthError")
ServiceTimeoutError = _k8s_only("ServiceTimeoutError")
VersionMismatchError = _k8s_only("VersionMismatchError")
NotebookError = _k8s_only("NotebookError")
DataStoreError = _k8s_only("DataStoreError")

EXCEPTION_REGISTRY = {
    cls.__name__: cls
    for cls in (
        ControllerRequestError, ImagePullError, KubernetesCredentialsError, PodContainerError,
        ResourceNotAvailableError, ServiceHealthError, ServiceTimeoutError, StartupError, PodTerminatedError,
        NotebookError, KnativeServiceConflictError, RsyncError, DataStoreError, VersionMismatchError,
        SecretNotFound, WorkerMembershipChanged,
    )
}

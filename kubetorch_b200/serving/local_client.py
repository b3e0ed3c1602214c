"""Client seam of a deployed module — same contract as the reference's HTTPClient
(kt/serving/http_client.py:1041-1108): call_method(endpoint, stream_logs, logging_config,
stream_metrics=None, body=None, headers=None, serialization="json") -> result, raising rebuilt
exceptions that carry .remote_traceback and .pod_name (http_client.py:154-175).

There is no HTTP hop on the local route: the "endpoint" names the callable/method and the call goes
straight into the supervisor.  JSON mode still normalises args/results through a JSON round trip
(tuples become lists, exactly what a caller of the reference observes) and enforces the
serialization allow-list with the reference's message.
"""
from __future__ import annotations

import asyncio
import json
import os
from typing import Optional

from .codec import MAGIC_CALL_KWARGS, SERIALIZATION_FORMATS, package_exception, rebuild_exception
from .supervisors import Request


# values a JSON round trip returns unchanged (exact types: a bool is not "an int" here, subclasses are not assumed)
_JSON_SCALARS = frozenset((str, int, float, bool, type(None)))


def _json_invariant(args, kwargs) -> bool:
    """True when json.loads(json.dumps(...)) would hand back equal objects: only top-level scalars."""
    for v in args:
        if type(v) not in _JSON_SCALARS:
            return False
    for v in kwargs.values():
        if type(v) not in _JSON_SCALARS:
            return False
    return True


def parse_endpoint(endpoint: str):
    """'local://<service>/<callable>[/<method>]' → (callable, method)."""
    path = endpoint.split("://", 1)[-1]
    parts = path.split("/")
    callable_name = parts[1] if len(parts) > 1 else parts[0]
    method = parts[2] if len(parts) > 2 else None
    return callable_name, method


class LocalClient:
    def __init__(self, supervisor, service_name: str, pod_name: Optional[str] = None):
        self.supervisor = supervisor
        self.service_name = service_name
        self.pod_name = pod_name or f"{service_name}-0"

    def _request_id(self, endpoint: str) -> str:
        # 10 hex digits, unique per call like the reference's sha256(endpoint, time)[:10] (http_client.py:311-349)
        return os.urandom(5).hex()

    def call_method(self, endpoint: str, stream_logs=None, logging_config=None, stream_metrics=None, body: dict = None,
                    headers: dict = None, serialization: str = "json"):
        if serialization not in SERIALIZATION_FORMATS:
            raise ValueError("Serialization must be 'json', 'pickle', or 'none'")
        callable_name, method = parse_endpoint(endpoint)
        body = dict(body or {})
        kwargs = dict(body.get("kwargs") or {})
        for magic in MAGIC_CALL_KWARGS:  # hoisted out of kwargs like _serialize_body does
            if magic in kwargs:
                body[magic] = kwargs.pop(magic)
        body["kwargs"] = kwargs
        if serialization == "json" and not _json_invariant(body.get("args", ()), kwargs):
            try:  # what httpx(json=...) would do to the args on the way out
                wire = json.loads(json.dumps({"args": body.get("args", []), "kwargs": kwargs}))
            except (TypeError, ValueError) as e:
                raise TypeError(f"Object of call arguments is not JSON serializable: {e}") from None
            body["args"], body["kwargs"] = wire["args"], wire["kwargs"]
        req_headers = {"X-Request-ID": self._request_id(endpoint), "X-Serialization": serialization}
        req_headers.update(headers or {})
        try:
            result = self.supervisor.call(Request(req_headers), callable_name, method, body)
        except BaseException as e:  # noqa: BLE001
            if hasattr(e, "remote_traceback"):
                raise  # already packaged by a rank process
            raise rebuild_exception(package_exception(e, pod_name=self.pod_name)) from None
        if serialization == "json" and type(result) not in _JSON_SCALARS:
            result = json.loads(json.dumps(result))
        return result

    async def call_method_async(self, endpoint: str, stream_logs=None, logging_config=None, stream_metrics=None,
                                body: dict = None, headers: dict = None, serialization: str = "json"):
        loop = asyncio.get_running_loop()
        return await loop.run_in_executor(
            None, lambda: self.call_method(endpoint, stream_logs, logging_config, stream_metrics, body, headers,
                                           serialization))

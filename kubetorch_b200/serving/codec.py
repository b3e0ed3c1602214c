"""Wire codecs and the error envelope of the remote-call path, byte-compatible with the reference:

    serialize_body         kt/serving/utils.py:730-749
    deserialize_response   kt/serving/utils.py:787-813
    parse_callable_params  kt/serving/http_server.py:1768-1822
    serialize_result       kt/serving/http_server.py:1825-1842
    package_exception      kt/serving/http_server.py:1478-1526
    raise_packaged         kt/serving/http_client.py:87-175 (CustomResponse.raise_for_status)

They are used (a) by the process-pool backends, where arguments really cross a process boundary,
and (b) to keep JSON/pickle/allow-list semantics observable from user code identical.  On the
B200 device route tensor leaves do NOT pass through these codecs: they are packed into HBM arenas
by ktb_pack (kubetorch_b200/device/ops.py) and only the small non-tensor header is pickled.
"""
from __future__ import annotations

import asyncio
import base64
import builtins
import json
import os
import pickle
import traceback
from typing import Any, Optional, Tuple

from ..exceptions import EXCEPTION_REGISTRY, SerializationError

MAGIC_CALL_KWARGS = ["workers", "restart_procs"]
DEFAULT_ALLOWED_SERIALIZATION = "json,pickle"
SERIALIZATION_FORMATS = ("json", "pickle", "none")


class HTTPException(Exception):
    """Same observable shape as fastapi.HTTPException: str(e) == f"{status_code}: {detail}"."""

    def __init__(self, status_code: int, detail: str):
        super().__init__(f"{status_code}: {detail}")
        self.status_code = status_code
        self.detail = detail


def serialize_body(body: Optional[dict], serialization: str) -> dict:
    if body is None:
        return {}
    kwargs = body.get("kwargs", {})
    for magic in MAGIC_CALL_KWARGS:
        if magic in kwargs:
            body[magic] = kwargs.pop(magic)
    if serialization == "pickle":
        payload = {"args": body.pop("args"), "kwargs": body.pop("kwargs")}
        body["data"] = base64.b64encode(pickle.dumps(payload)).decode("utf-8")
    return body


def deserialize_response(response_json: Any, serialization: str) -> Any:
    if serialization != "pickle":
        return response_json

    def unwrap(item):
        if isinstance(item, dict) and "data" in item:
            return pickle.loads(base64.b64decode(item["data"].encode("utf-8")))
        return item

    if isinstance(response_json, list):
        return [unwrap(r) for r in response_json]
    return unwrap(response_json)


def check_allowed(serialization: str, allowed: Optional[str] = None) -> None:
    allowed_list = (allowed if allowed is not None
                    else os.getenv("KT_ALLOWED_SERIALIZATION", DEFAULT_ALLOWED_SERIALIZATION)).split(",")
    if serialization not in allowed_list:
        raise HTTPException(400, f"Serialization format '{serialization}' not allowed. Allowed formats: {allowed_list}")


def parse_callable_params(params: Optional[dict], serialization: str, allowed: Optional[str] = None) -> Tuple[list, dict]:
    check_allowed(serialization, allowed)
    args, kwargs = [], {}
    if params:
        if serialization == "pickle":
            if isinstance(params, dict) and "data" in params:
                params.update(pickle.loads(base64.b64decode(params.pop("data").encode("utf-8"))))
            elif isinstance(params, str):
                params = pickle.loads(base64.b64decode(params.encode("utf-8")))
        args = params.get("args", [])
        kwargs = params.get("kwargs", {})
    return args, kwargs


def serialize_result(result: Any, serialization: str) -> Any:
    if serialization == "pickle":
        try:
            return {"data": base64.b64encode(pickle.dumps(result)).decode("utf-8")}
        except Exception as e:  # noqa: BLE001
            raise SerializationError(f"Result could not be serialized with pickle: {e}")
    if serialization == "json":
        try:
            json.dumps(result)
        except (TypeError, ValueError) as e:
            raise SerializationError(f"Result could not be serialized to JSON: {e}")
    return result


def status_code_for(exc: BaseException) -> int:
    import concurrent.futures

    if hasattr(exc, "status_code"):
        return exc.status_code
    if isinstance(exc, (TypeError, AssertionError)):
        return 422
    if isinstance(exc, (ValueError, UnicodeError)):
        return 400
    if isinstance(exc, (KeyError, FileNotFoundError)):
        return 404
    if isinstance(exc, PermissionError):
        return 403
    if isinstance(exc, (MemoryError, OSError)):
        return 500
    if isinstance(exc, NotImplementedError):
        return 501
    if isinstance(exc, (asyncio.TimeoutError, concurrent.futures.TimeoutError)):
        return 504
    return 500


def package_exception(exc: BaseException, pod_name: Optional[str] = None) -> dict:
    """Error envelope {error_type, message, traceback, pod_name, state, status_code}."""
    state = None
    if hasattr(exc, "__getstate__"):
        try:
            state = exc.__getstate__()
            json.dumps(state)
        except Exception:  # noqa: BLE001
            state = None
    return {
        "error_type": exc.__class__.__name__,
        "message": str(exc),
        "traceback": "".join(traceback.format_exception(type(exc), exc, exc.__traceback__)),
        "pod_name": pod_name or os.getenv("POD_NAME", "unknown"),
        "state": state,
        "status_code": status_code_for(exc),
    }


def is_error_envelope(obj: Any) -> bool:
    return isinstance(obj, dict) and all(k in obj for k in ("error_type", "message", "traceback", "pod_name"))


def rebuild_exception(envelope: dict) -> BaseException:
    error_type, message = envelope["error_type"], envelope.get("message", "")
    state = envelope.get("state") or {}
    cls = getattr(builtins, error_type, None)
    if not (isinstance(cls, type) and issubclass(cls, BaseException)):
        cls = EXCEPTION_REGISTRY.get(error_type)
    exc = None
    if cls is not None:
        try:
            exc = cls.from_dict(state) if (state and hasattr(cls, "from_dict")) else cls(message)
        except Exception:  # noqa: BLE001
            exc = None
    if exc is None:
        exc = type(error_type, (Exception,), {})(message)
    exc.remote_traceback = envelope["traceback"]
    exc.pod_name = envelope["pod_name"]
    if "status_code" in envelope and not hasattr(exc, "status_code"):
        exc.status_code = envelope["status_code"]

    class RemoteException(exc.__class__):
        def __str__(self):
            try:
                cleaned = self.remote_traceback.encode().decode("unicode_escape")
            except Exception:  # noqa: BLE001
                cleaned = self.remote_traceback
            return f"{super().__str__()}\n\n{cleaned}"

    RemoteException.__name__ = exc.__class__.__name__
    RemoteException.__qualname__ = exc.__class__.__qualname__
    wrapped = RemoteException.__new__(RemoteException)
    wrapped.__dict__.update(exc.__dict__)
    wrapped.args = (str(exc),)
    return wrapped


def raise_packaged(envelope: dict):
    raise rebuild_exception(envelope)

"""Per-call / per-compute configuration dataclasses with the reference's field names
(kt/globals.py:39-120) so user code constructing them keeps working.  Log/metric streaming has no
Loki/Prometheus behind it on the local route: the values are accepted and ignored."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Literal


def _log_level() -> str:
    level = os.getenv("KT_LOG_LEVEL", "INFO").lower()
    valid = ("debug", "info", "warning", "error", "critical")
    if level not in valid:
        raise ValueError(f"Invalid KT_LOG_LEVEL environment variable: '{level}'. Must be one of: {valid}")
    return level


@dataclass
class MetricsConfig:
    interval: int = 30
    scope: Literal["pod", "resource"] = "resource"


@dataclass
class LoggingConfig:
    stream_logs: bool = None
    level: Literal["debug", "info", "warning", "error", "critical"] = field(default_factory=_log_level)
    include_system_logs: bool = False
    include_events: bool = True
    grace_period: float = 2.0
    include_name: bool = True
    poll_timeout: float = 1.0
    grace_poll_timeout: float = 0.5
    shutdown_grace_period: float = 0


@dataclass
class DebugConfig:
    mode: Literal["pdb", "pdb-ui"] = "pdb"
    port: int = 5678

    def to_dict(self):
        return {"mode": self.mode, "port": self.port}


class KubetorchConfig:
    """Minimal stand-in for kt.config (kt/config.py:13-383) with the reference's precedence:
    explicit setter > KT_* env > ~/.kt/config.yaml > default."""

    CONFIG_FILE = "~/.kt/config.yaml"
    _DEFAULTS = {"stream_logs": True, "stream_metrics": False, "namespace": "default", "username": None,
                 "backend": "auto"}

    def __init__(self):
        self._explicit = {}
        self._file_cache = None

    def _from_file(self) -> dict:
        if self._file_cache is None:
            path = os.path.expanduser(self.CONFIG_FILE)
            data = {}
            if os.path.exists(path):
                try:
                    import yaml

                    with open(path) as f:
                        data = yaml.safe_load(f) or {}
                except Exception:  # noqa: BLE001 - an unreadable config file is ignored, as in the reference
                    data = {}
            self._file_cache = data if isinstance(data, dict) else {}
        return self._file_cache

    def refresh(self):
        """Forget the cached file contents (tests; the reference caches per process too)."""
        self._file_cache = None

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name in self._explicit:
            return self._explicit[name]
        env = os.getenv(f"KT_{name.upper()}")
        if env is not None:
            return {"true": True, "false": False}.get(env.lower(), env)
        if name in self._from_file():
            return self._file_cache[name]
        if name in self._DEFAULTS:
            return self._DEFAULTS[name]
        raise AttributeError(f"kt.config has no setting '{name}'")

    def set(self, name, value):
        self._explicit[name] = value

    def __setattr__(self, name, value):
        if name.startswith("_"):
            object.__setattr__(self, name, value)
        else:
            self._explicit[name] = value


config = KubetorchConfig()

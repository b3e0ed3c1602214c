"""kt.distributed helpers (kt/distributed/utils.py:19-129): on one box the "pods" are local ranks."""
from __future__ import annotations

import os
from typing import List, Optional


def local_pod_ips(workers: int) -> List[str]:
    """Loopback addresses standing in for pod IPs: 127.0.0.1 … 127.0.0.<workers> (all route locally)."""
    return [f"127.0.0.{k + 1}" for k in range(max(1, int(workers)))]


def pod_ips(quorum_workers: Optional[int] = None, quorum_timeout: Optional[int] = None) -> List[str]:
    """Worker addresses of the current service: POD_IPS (set per rank) or LOCAL_IPS, else 127.0.0.1."""
    for var in ("POD_IPS", "LOCAL_IPS"):
        val = os.environ.get(var)
        if val:
            return [ip for ip in val.split(",") if ip]
    return local_pod_ips(quorum_workers or 1)

#!/usr/bin/env python
"""Recipe: place the UNMODIFIED reference (run-house/kubetorch @ 96fac95, python_client v0.5.0) under
baseline/_ref/ so `bench.py --impl reference` can run it on the GPU box, where /root/reference does not exist.

  * `pip install --no-index --no-build-isolation --target baseline/_ref /root/reference/python_client` cannot work
    here: the build backend is poetry-core (python_client/pyproject.toml:97-99), which is in neither the image nor
    /opt/wheelhouse ("ModuleNotFoundError: No module named 'poetry'").  The package is pure Python with no build
    step, so installing it IS copying the package directory — which is what this script does, byte for byte
    (baseline/_ref/MANIFEST.json records a sha256 per file so the copy can be audited against the reference).
  * the one import the image lacks, `websocket` (websocket-client; imported at kubetorch/data_store/
    websocket_tunnel.py:8, only USED at :161 for the rsync tunnel, never on this path), gets a 3-line stub module.

baseline/_ref/ is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored,
so it travels to the GPU box with the snapshot.  Runs only where /root/reference is mounted; `__graft_entry__.build()`
calls it there.  Usage: python baseline/install_reference.py
"""
import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/python_client/kubetorch"
DST = os.path.join(REPO, "baseline", "_ref")

STUB = (
    '"""Stub for the absent websocket-client package (see baseline/install_reference.py)."""\n'
    "class WebSocketException(Exception): pass\n"
    "ABNF = type('ABNF', (), {'OPCODE_BINARY': 2, 'OPCODE_TEXT': 1})\n"
    "def create_connection(*a, **k): raise RuntimeError('websocket-client is not installed (stub)')\n"
)


def install(quiet: bool = False) -> bool:
    if not os.path.isdir(SRC):
        if not quiet:
            print(f"{SRC} not found: the reference is only mounted in the authoring container", file=sys.stderr)
        return False
    pkg = os.path.join(DST, "kubetorch")
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(SRC, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "websocket.py"), "w") as f:
        f.write(STUB)
    manifest = {}
    for root, _, files in os.walk(pkg):
        for name in sorted(files):
            path = os.path.join(root, name)
            with open(path, "rb") as fh:
                manifest[os.path.relpath(path, DST)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "reference": "run-house/kubetorch @ 96fac95 (python_client v0.5.0), unmodified",
                   "files": manifest}, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"installed {len(manifest)} files into {pkg}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)

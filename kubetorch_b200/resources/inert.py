"""Call-site stand-ins for the cluster-side resources a kubetorch program usually names next to its compute:
kt.Image / kt.images.*, kt.Secret / kt.secret, kt.Volume (kt/resources/images/image.py:6-560, images.py:1-47,
secrets/secret.py:9, secrets/secret_factory.py:8, volumes/volume.py:17).

They are OUT OF SCOPE on the local-B200 route (no container is built, no secret or volume is mounted — DESIGN.md §6): the
objects record what was asked so programs written against the reference construct and pass them unchanged.  The one
setting with a local meaning is Image.set_env_vars: those variables are exported to the rank processes."""
from __future__ import annotations

from typing import Dict, List, Optional, Union


class Image:
    def __init__(self, name: str = None, image_id: str = None, python_path: str = None, install_cmd: str = None):
        self.name = name
        self.image_id = image_id
        self.python_path = python_path
        self.install_cmd = install_cmd
        self.env_vars: Dict[str, str] = {}
        self.steps: List[tuple] = []          # what a container build would have run, for inspection

    def from_docker(self, image_id: str):
        self.image_id = image_id
        return self

    @classmethod
    def from_dockerfile(cls, dockerfile_path: str, name: str = None) -> "Image":
        img = cls(name=name)
        img.steps.append(("dockerfile", dockerfile_path))
        return img

    def pip_install(self, reqs: Union[str, List[str]], force: bool = False):
        self.steps.append(("pip_install", [reqs] if isinstance(reqs, str) else list(reqs)))
        return self

    def set_env_vars(self, env_vars: Dict):
        self.env_vars.update({str(k): str(v) for k, v in env_vars.items()})
        return self

    def sync_package(self, package: str, force: bool = False):
        self.steps.append(("sync_package", package))
        return self

    def run_bash(self, command: str, force: bool = False):
        self.steps.append(("run_bash", command))
        return self

    def copy(self, source: str, dest: str = None, contents: bool = False, force: bool = False):
        self.steps.append(("copy", (source, dest)))
        return self

    def rsync(self, source: str, dest: str = None, contents: bool = False, filter_options: str = None,
              force: bool = False):
        self.steps.append(("rsync", (source, dest)))
        return self

    def __repr__(self):
        return f"Image(name={self.name!r}, image_id={self.image_id!r}, steps={len(self.steps)})"


class _Images:
    """kt.images: debian() / ubuntu() / python(v) / ray(v) / pytorch(v) and the capitalised aliases."""

    @staticmethod
    def debian() -> Image:
        return Image(name="debian", image_id="kubetorch-server-minimal")

    @staticmethod
    def ubuntu() -> Image:
        return Image(name="ubuntu", image_id="kubetorch-ubuntu-minimal")

    @staticmethod
    def python(version: str) -> Image:
        return Image(name=f"python{version.replace('.', '')}", image_id=f"python:{version}-slim")

    @staticmethod
    def ray(version: str = "latest") -> Image:
        return Image(name=f"ray{version if version != 'latest' else ''}".strip(), image_id=f"rayproject/ray:{version}")

    @staticmethod
    def pytorch(version: str = "23.12-py3") -> Image:
        return Image(name=f"pytorch{version.replace('.', '').replace('-', '')}",
                     image_id=f"nvcr.io/nvidia/pytorch:{version}")

    Debian = staticmethod(lambda: _Images.debian())
    Ubuntu = staticmethod(lambda: _Images.ubuntu())
    Ray = staticmethod(lambda: _Images.ray("latest"))
    Pytorch2312 = staticmethod(lambda: _Images.pytorch("23.12-py3"))
    Python310 = staticmethod(lambda: _Images.python("3.10"))
    Python311 = staticmethod(lambda: _Images.python("3.11"))
    Python312 = staticmethod(lambda: _Images.python("3.12"))


images = _Images()


class Secret:
    def __init__(self, name: str = None, provider: str = None, values: Dict = None, path: str = None,
                 env_vars: Dict = None, override: bool = False, **kwargs):
        self.name, self.provider, self.values, self.path = name, provider, values, path
        self.env_vars, self.override = env_vars, override


def secret(name: Optional[str] = None, provider: Optional[str] = None, path: str = None, env_vars: Dict = None,
           override: bool = False, **kwargs) -> Secret:
    if not (name or provider):
        raise ValueError("Either name or provider must be provided")
    return Secret(name=name, provider=provider, path=path, env_vars=env_vars, override=override)


class Volume:
    def __init__(self, name: str = None, size: str = None, mount_path: str = None, storage_class: str = None,
                 access_mode: str = None, namespace: str = None, **kwargs):
        self.name, self.size, self.mount_path = name, size, mount_path
        self.storage_class, self.access_mode, self.namespace = storage_class, access_mode, namespace

"""Import shim that lets the UNMODIFIED reference (/root/reference/src/super_gradients) run in the build
container, where omegaconf / hydra / torchmetrics / onnx / ... are not installed.

TEST INFRASTRUCTURE ONLY.  It is used by ``tests/golden/make_goldens.py`` (run once, in the build
container, output committed under tests/golden/) and by optional ``-m "not gpu"`` tests that skip when
/root/reference is absent.  Nothing on the product path imports this file, and nothing on the GPU box
can (the reference tree does not exist there).

How it works:
  1. a ``sys.meta_path`` finder serves permissive stub modules for the missing third-party packages;
  2. ``super_gradients`` is pre-seeded as an empty package rooted at the reference source tree so that
     sub-modules import lazily, the yaml arch-param loader is replaced by a PyYAML one, and then the real
     ``__init__`` is executed.
"""
import copy
import importlib
import importlib.abc
import importlib.machinery
import os
import re
import sys
import types

REF_ROOT = os.environ.get("SG_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")
SG_DIR = os.path.join(REF_SRC, "super_gradients")

_MISSING = (
    "omegaconf", "hydra", "torchmetrics", "onnx", "onnxruntime", "onnxsim", "albumentations", "treelib", "termcolor",
    "stringcase", "rapidfuzz", "json_tricks", "data_gradients", "pycocotools", "boto3", "botocore", "matplotlib",
    "imagesize", "coverage", "deprecated", "wandb", "clearml", "dagshub", "mlflow", "pytorch_quantization", "onnx_graphsurgeon",
    "tensorrt", "pip_tools", "sphinx", "pyparsing", "einops_exts", "imutils", "lightning_utilities",
)


def available() -> bool:
    return os.path.isdir(SG_DIR)


class _Meta(type):
    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _make_dummy(f"{cls.__name__}.{name}")

    def __call__(cls, *a, **k):
        # used as decorator: @stub(...) / @stub
        if len(a) == 1 and not k and (isinstance(a[0], type) or callable(a[0])) and not isinstance(a[0], (str, int, float)):
            return a[0]
        return type.__call__(cls)

    def __instancecheck__(cls, inst):
        return False

    def __subclasscheck__(cls, sub):
        return False

    def __iter__(cls):
        return iter(())

    def __getitem__(cls, item):
        return cls

    def __or__(cls, other):
        return cls

    def __ror__(cls, other):
        return cls


def _make_dummy(name):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and not k and callable(a[0]):
            return a[0]
        return self

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _make_dummy(item)

    return _Meta(name.split(".")[-1], (), {"__init__": __init__, "__call__": __call__, "__getattr__": __getattr__, "__module__": "stub"})


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        d = _make_dummy(f"{self.__name__}.{name}")
        setattr(self, name, d)
        return d


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in _MISSING:
            try:
                # genuinely installed? then leave it alone
                for f in sys.meta_path:
                    if f is self:
                        continue
                    spec = f.find_spec(fullname, path, target) if hasattr(f, "find_spec") else None
                    if spec is not None:
                        return None
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_FLOAT_RE = re.compile(r"^[+-]?\d+(\.\d*)?[eE][+-]?\d+$")


def _fix(o):
    if isinstance(o, dict):
        return {k: _fix(v) for k, v in o.items() if k != "_convert_"}
    if isinstance(o, list):
        return [_fix(v) for v in o]
    if isinstance(o, str) and _FLOAT_RE.match(o):
        return float(o)
    return o


def _deep_merge(a, b):
    out = dict(a)
    for k, v in b.items():
        if k in out and isinstance(out[k], dict) and isinstance(v, dict):
            out[k] = _deep_merge(out[k], v)
        else:
            out[k] = v
    return out


def load_arch_params_yaml(config_name: str, recipes_dir_path=None, overriding_params=None):
    import yaml

    base = os.path.join(SG_DIR, "recipes", "arch_params")
    name = config_name if config_name.endswith(".yaml") else config_name + ".yaml"
    with open(os.path.join(base, name)) as f:
        cfg = yaml.safe_load(f) or {}
    merged = {}
    for d in cfg.pop("defaults", []) or []:
        if isinstance(d, str) and d != "_self_":
            merged = _deep_merge(merged, load_arch_params_yaml(d))
    merged = _deep_merge(merged, cfg)
    if overriding_params:
        merged = _deep_merge(merged, dict(overriding_params))
    return _fix(merged)


_installed = None


def install():
    """Returns the imported reference ``super_gradients`` package (cached)."""
    global _installed
    if _installed is not None:
        return _installed
    if not available():
        raise RuntimeError(f"reference tree not found at {SG_DIR}")
    os.environ.setdefault("CRASH_HANDLER", "FALSE")
    os.environ.setdefault("CONSOLE_LOG_LEVEL", "ERROR")
    os.environ.setdefault("FILE_LOG_LEVEL", "ERROR")
    os.environ.setdefault("SUPER_GRADIENTS_LOG_DIR", "/tmp/sg_logs")
    sys.meta_path.append(_StubFinder())
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    pkg = types.ModuleType("super_gradients")
    pkg.__path__ = [SG_DIR]
    pkg.__file__ = os.path.join(SG_DIR, "__init__.py")
    pkg.is_distributed = lambda: False
    sys.modules["super_gradients"] = pkg
    cfg_utils = importlib.import_module("super_gradients.common.environment.cfg_utils")
    cfg_utils.load_arch_params = load_arch_params_yaml
    apf = importlib.import_module("super_gradients.training.models.arch_params_factory")
    apf.load_arch_params = load_arch_params_yaml
    apf.hydra.utils.instantiate = copy.deepcopy
    with open(pkg.__file__) as f:
        code = compile(f.read(), pkg.__file__, "exec")
    exec(code, pkg.__dict__)
    pkg.is_distributed = lambda: False
    _installed = pkg
    return pkg


if __name__ == "__main__":
    sg = install()
    import torch

    from super_gradients.training import models

    m = models.get("yolo_nas_s", num_classes=80)
    m.eval()
    with torch.no_grad():
        out = m(torch.randn(1, 3, 64, 64))
    print("ok", out[0][0].shape, out[0][1].shape, sum(p.numel() for p in m.parameters()))

"""get_arch_params: yaml -> dict (reference: training/models/arch_params_factory.py + hydra compose; here a plain PyYAML
loader with `defaults:` deep-merge -- the recipes' config system itself is out of scope)."""
import copy
import os
import re

import yaml

_RECIPES = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "recipes", "arch_params")
_FLOAT = re.compile(r"^[+-]?\d+(\.\d*)?[eE][+-]?\d+$")


def _fix(o):
    if isinstance(o, dict):
        return {k: _fix(v) for k, v in o.items() if k != "_convert_"}
    if isinstance(o, list):
        return [_fix(v) for v in o]
    if isinstance(o, str) and _FLOAT.match(o):
        return float(o)  # PyYAML reads `1e-3` as a string
    return o


def _merge(a, b):
    out = dict(a)
    for k, v in b.items():
        out[k] = _merge(out[k], v) if k in out and isinstance(out[k], dict) and isinstance(v, dict) else v
    return out


def get_arch_params(config_name: str, overriding_params: dict = None, recipes_dir_path: str = None) -> dict:
    base = recipes_dir_path or _RECIPES
    name = config_name if config_name.endswith(".yaml") else config_name + ".yaml"
    path = os.path.join(base, name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"arch params `{config_name}` not found under {base}")
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    merged = {}
    for d in cfg.pop("defaults", []) or []:
        if isinstance(d, str) and d != "_self_":
            merged = _merge(merged, get_arch_params(d, recipes_dir_path=recipes_dir_path))
    merged = _merge(merged, cfg)
    if overriding_params:
        merged = _merge(merged, dict(overriding_params))
    return copy.deepcopy(_fix(merged))

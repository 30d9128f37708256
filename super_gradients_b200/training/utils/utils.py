"""HpmStruct / get_param (reference: training/utils/utils.py)."""
import copy
from collections.abc import Mapping


class HpmStruct:
    def __init__(self, **entries):
        self.__dict__.update(entries)
        self.schema = None

    def set_schema(self, schema: dict):
        self.schema = schema

    def override(self, **entries):
        recursive_override(self.__dict__, entries)

    def to_dict(self, include_schema=True) -> dict:
        out = dict(self.__dict__)
        if not include_schema:
            out.pop("schema", None)
        return out

    def validate(self):
        pass


def recursive_override(base: dict, extension: dict):
    for k, v in extension.items():
        if k in base and isinstance(base[k], Mapping) and isinstance(v, Mapping):
            recursive_override(base[k], v)
        else:
            base[k] = copy.deepcopy(v)


def get_param(params, name, default_val=None):
    if isinstance(params, Mapping):
        if name in params:
            v = params[name]
            if isinstance(v, Mapping) and v and default_val is not None and isinstance(default_val, Mapping):
                return {**default_val, **v}
            return v
        return default_val
    if hasattr(params, name):
        return getattr(params, name)
    return default_val

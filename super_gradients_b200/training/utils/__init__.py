from .utils import HpmStruct, get_param  # noqa: F401

"""width_multiplier / autopad (reference: modules/utils.py:63-74)."""
import math


def width_multiplier(original, factor, divisor: int = None):
    if divisor is None:
        return int(original * factor)
    return math.ceil(int(original * factor) / divisor) * divisor


def autopad(kernel, padding=None):
    if padding is None:
        padding = kernel // 2 if isinstance(kernel, int) else [x // 2 for x in kernel]
    return padding

"""``nengo.utils.numpy`` names used by ``sspspace.py:9``."""
import numpy as np

maxint = np.iinfo(np.int32).max


def is_integer(obj):
    return isinstance(obj, (int, np.integer)) and not isinstance(obj, bool)


def is_number(obj):
    return isinstance(obj, (int, float, np.number)) and not isinstance(obj, bool)


def is_iterable(obj):
    if isinstance(obj, np.ndarray):
        return obj.ndim > 0
    try:
        iter(obj)
        return True
    except TypeError:
        return False


def is_array_like(obj):
    return isinstance(obj, (np.ndarray, list, tuple))

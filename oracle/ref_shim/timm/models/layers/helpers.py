"""timm 0.4.5 `to_2tuple` restated (used by reference src/models/passt/passt.py:19)."""
import collections.abc


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return x
    return (x, x)

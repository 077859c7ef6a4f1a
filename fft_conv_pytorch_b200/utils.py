"""Argument normalisation shared by the functional and module API."""
from typing import Iterable, Tuple, Union


def to_ntuple(val: Union[int, Iterable[int]], n: int) -> Tuple[int, ...]:
    """Same contract as the reference's ``to_ntuple`` (reference fft_conv_pytorch/utils.py:4-20): an int becomes an
    n-tuple, an iterable must already have n entries (``ValueError`` otherwise — which also rejects string paddings
    such as ``'same'``, as the reference does)."""
    if isinstance(val, Iterable):
        out = tuple(val)
        if len(out) != n:
            raise ValueError(f"Cannot cast tuple of length {len(out)} to length {n}.")
        return out
    return n * (val,)

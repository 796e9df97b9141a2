"""Arithmetic mode of the conv stacks.

'fp32' : activations fp32, convolutions on the CUDA cores in fp32 — parity grade (1e-3 vs the reference).
'bf16' : activations bf16 (NHWC), convolutions on the tcgen05 tensor cores with fp32 accumulation,
         BatchNorm statistics / soft-max / render / loss / Adam in fp32 — throughput mode.
"""
from contextlib import contextmanager

_precision = 'fp32'


def get_precision() -> str:
    return _precision


def set_precision(p: str) -> None:
    global _precision
    if p not in ('fp32', 'bf16'):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _precision = p


@contextmanager
def precision(p: str):
    old = get_precision()
    set_precision(p)
    try:
        yield
    finally:
        set_precision(old)

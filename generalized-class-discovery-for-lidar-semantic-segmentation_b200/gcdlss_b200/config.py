"""Global math-mode switch of the sparse-conv path.

``fp32``: every kernel computes and stores fp32 (SIMT FMA convolution); meets the 1e-4 parity
bound against the oracle.  ``bf16``: features are stored in bf16 between layers and the
convolutions run on tcgen05 tensor cores with fp32 accumulation in TMEM; parameters, BN
statistics, weight gradients and everything a caller sees through ``SparseTensor.F`` stay fp32.
"""
import os

_MODES = ("fp32", "bf16")
_state = {"mode": os.environ.get("GCDLSS_MATH", "fp32")}
if _state["mode"] not in _MODES:
    raise ValueError(f"GCDLSS_MATH must be one of {_MODES}")


def set_math_mode(mode: str) -> None:
    if mode not in _MODES:
        raise ValueError(f"math mode must be one of {_MODES}")
    _state["mode"] = mode


def get_math_mode() -> str:
    return _state["mode"]

"""ctypes binding of the C-ABI CUDA library ``libholocron_b200.so`` (declared in ``include/holocron_b200.h``).

The library is the product: there is no Python/PyTorch fallback for any op it implements. ``lib()`` raises
``RuntimeError`` when the shared object is missing or a kernel launch reports an error.
"""
import ctypes
import os
from pathlib import Path
from typing import Optional

import torch

_LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libholocron_b200.so"
_lib: Optional[ctypes.CDLL] = None

DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


class HolocronB200Error(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    """Returns the loaded C-ABI library, loading it on first use. Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise HolocronB200Error(
                f"{_LIB_PATH} is missing: build it with `python -m holocron_b200.csrc.build` "
                "(there is no CPU / PyTorch fallback for the holocron_b200 kernels)"
            )
        _lib = ctypes.CDLL(os.fspath(_LIB_PATH))
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise HolocronB200Error(f"{what} failed with CUDA error code {rc}")


def ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(t: torch.Tensor) -> int:
    try:
        return DTYPE_CODE[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}: expected float32, bfloat16 or float16") from None


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise HolocronB200Error(
                "holocron_b200 ops run on CUDA tensors only (no CPU fallback); "
                "use the reference implementation / oracle for CPU tensors"
            )

"""freeimpala-b200: B200-native learner hot path for freeimpala (sm_100a CUDA behind a C ABI).

    SharedBuffer, Learner, ModelManager   host-side mirrors of the reference interfaces
    ops                                   stream-ordered launches of the individual kernels
    build                                 compiles freeimpala_b200/_build/libfreeimpala_b200.so

There is no CPU / PyTorch fallback: without the compiled library every entry point raises.
"""
from . import _lib
from ._lib import FiError, load as load_library
from .learner import Learner, Model, ModelManager
from .shared_buffer import Batch, SharedBuffer, ELEMENT_SIZE
from . import ops

__all__ = ["Learner", "Model", "ModelManager", "SharedBuffer", "Batch", "ELEMENT_SIZE", "ops", "FiError",
           "load_library", "kernel_launch_count", "version"]


def kernel_launch_count() -> int:
    return load_library().fi_kernel_launch_count()


def version() -> str:
    return load_library().fi_version().decode()


def prof_enable(on: bool = True) -> None:
    load_library().fi_prof_enable(int(on))


def prof_collect() -> dict:
    """{kernel name: {"launches", "total_ms", "work", "unit"}} for the launches since the last call."""
    import ctypes as C
    arr = (_lib.FiProfEntry * 64)()
    n = min(load_library().fi_prof_collect(arr, 64), 64)
    return {arr[i].name.decode(): {"launches": int(arr[i].launches), "total_ms": float(arr[i].total_ms),
                                   "work": float(arr[i].work), "unit": "flops" if arr[i].unit else "bytes"}
            for i in range(n)}

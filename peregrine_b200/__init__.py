"""peregrine_b200 — B200-native SHIMMER index + read-to-read overlap engine.

The product is ``libpgb200.so`` (CUDA kernels for sm_100a behind the C ABI in ``include/pgb200.h``) and the drop-in command
line tools ``bin/shmr_{mkseqdb,index,overlap,dedup,map}``.  This package is the thin Python host mirror:

* :mod:`peregrine_b200.engine`    ctypes binding of the stage-level API (``Engine``)
* :mod:`peregrine_b200.formats`   numpy views of the reference's on-disk formats (Appendix B of SURVEY.md)
* :mod:`peregrine_b200.shimmer4py` cffi ``lib`` object with the same cdef as ``peregrine._shimmer4py``
  (py/peregrine/build_shimmer4py.py:8-84 of the reference)
* :mod:`peregrine_b200.utils`     the SHIMMER helpers of ``peregrine.utils`` (same names and signatures) on that ``lib``
* :mod:`peregrine_b200.multigpu`  one-process-per-GPU driver (torch.distributed / NCCL) of the sharded job

There is no CPU implementation here: importing works anywhere, computing needs a CUDA device.
"""
from .engine import Engine, lib_path, load_library, NoDeviceError  # noqa: F401
from . import formats  # noqa: F401

__all__ = ["Engine", "lib_path", "load_library", "NoDeviceError", "formats"]

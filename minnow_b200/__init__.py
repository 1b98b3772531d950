"""minnow_b200 -- the block encode/decode hot path of phil-mansfield/minnow on
NVIDIA B200 (sm_100a), behind the C ABI of include/minnow_cuda.h.

The package is a thin host-side mirror of the reference's interfaces over
libminnow_b200.so.  There is no CPU implementation here: importing works
anywhere, but creating a Context needs the built library and a CUDA device.
"""
from .capi import (Context, Pipe, MinnowError, library_path, load_library, precision_needed,  # noqa: F401
                   array_bytes, float_group_pixels, jitter_hash32, FloatDesc, Jitter,
                   JITTER_CENTER, JITTER_HASH, JITTER_STREAM)

from . import minnow, minh, minp, shard  # noqa: F401  (host-side mirrors of the reference's packages)

__all__ = ["Context", "Pipe", "MinnowError", "library_path", "load_library", "precision_needed", "array_bytes",
           "float_group_pixels", "jitter_hash32", "FloatDesc", "Jitter", "JITTER_CENTER", "JITTER_HASH",
           "JITTER_STREAM"]

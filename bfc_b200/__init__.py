"""bfc_b200 -- host-side Python mirror of the B200 count/correct engine.

The product is `lib/libbfc_b200.so` (hand-written sm_100a CUDA behind the reference's
own C API, see include/*.h) and the `lib/bfc` command line.  This package only binds
that C ABI with ctypes so tests and bench.py can drive it; it contains no compute and
no CPU fallback: importing works anywhere, every call needs a B200.
"""
from .api import (Opt, Batch, Stats, Engine, lib, lib_path, make_opt, opt_by_size, build, BfcError,
                  records_to_batch)

__all__ = ["Opt", "Batch", "Stats", "Engine", "lib", "lib_path", "make_opt", "opt_by_size", "build",
           "BfcError", "records_to_batch"]

"""xelis_he_b200 -- B200 (sm_100a) batch verifier for XELIS-HE confidential transactions.

The product is the C-ABI library `libxhe_cuda.so` (include/xhe.h); this package is the thin ctypes binding used by the
tests and bench.py.  There is no CPU fallback: importing works anywhere (so the export table can be checked), but every
compute call needs a CUDA device and raises `XheError` otherwise.  Nothing here imports `oracle/`.
"""
from ._lib import XheError, Ctx, DeviceLedger, Ecdlp, lib_path, load_library, ERR_NAMES  # noqa: F401

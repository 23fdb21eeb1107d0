"""ckfft_b200 -- B200-native (sm_100a) implementation of the ckfft transform hot path.

The product is the C-ABI shared library `ckfft_b200/lib/libckfft_b200.so` (headers in `include/ckfft/`);
this package is the thin Python mirror of the reference interface used by the tests and bench.
"""
from .api import BOTH, FORWARD, INVERSE, CkFftError, Context, MultiContext, get_plan, kernel_launches, last_error  # noqa: F401

"""fcwdm: B200 (sm_100a) kernels for the fast-cwdm hot path behind a C-ABI (include/fcwdm.h).

`native` is the ctypes binding, `ops` the torch-tensor wrappers.  The drop-in packages `DWT_IDWT` and
`guided_diffusion` next to this one mirror the reference's import paths and call into `ops`."""
from . import native  # noqa: F401
from .native import FcwdmError  # noqa: F401

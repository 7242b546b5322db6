"""mtl_b200: ctypes binding of libmtl_b200.so (hand-written sm_100a kernels + host engine) and the
flat parameter arena the reference-compatible ``models`` / ``modules`` / ``trainer`` packages sit on.

There is no CPU fallback: every compute entry point raises if the library or a CUDA device is
missing."""
from .lib import MtlError, get_lib, library_path  # noqa: F401
from .spec import ModelSpec, param_specs  # noqa: F401
from .session import Batch, MetaStepper, Session  # noqa: F401
from .lm_session import LmSession, LmSpec, lm_param_specs  # noqa: F401

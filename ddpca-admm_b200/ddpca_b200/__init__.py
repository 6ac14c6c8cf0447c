"""ddpca_b200 -- host-side mirror of the reference's solver surface over libddpca_b200.so.

`MGPIS` below keeps the names, argument meaning and conventions of the reference's
`class MGPIS` (MGPIS.h:8-38); every compute call goes through the C ABI declared in
include/ddpca_b200.h into hand-written sm_100a kernels.  There is no CPU fallback:
without the built library or without a CUDA device the calls raise.
"""
from .lib import DdpcaError, load_library, device_count, library_path  # noqa: F401
from .mgpis import MGPIS, SMOOTH_LEX, SMOOTH_MC, Plan, KERNEL_CLASSES, setup_dryrun  # noqa: F401
from .mcontact import MCONTACT, DIRE_SOLV, VECT_MEDI_OSCI, gamma_project  # noqa: F401
from . import ddpk  # noqa: F401

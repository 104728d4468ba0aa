"""ferreus_rbf_rs_b200 — B200-native (sm_100a) drop-in for the BBFMM matvec + RBF solve hot path of
graphic-goose/ferreus_rbf_rs.  The package is a thin host-side mirror of the reference's Python
modules (``ferreus_bbfmm``, ``ferreus_rbf``) over the C ABI in ``include/ferreus_b200.h``."""
from . import _lib  # noqa: F401
from .bbfmm import (FmmKernelType, FmmParams, FmmTree, KernelParams, M2LCompressionType,  # noqa: F401
                    SpheroidalOrder)
from . import config, interpolant_config, progress  # noqa: F401,E402
from .rbf import Coefficients, RBFInterpolator  # noqa: F401,E402

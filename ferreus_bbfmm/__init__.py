"""Importable stand-in for the reference's PyO3 module ``ferreus_bbfmm``
(py_ferreus_bbfmm/src/lib.rs:15-24, stubs py_ferreus_bbfmm/ferreus_bbfmm/ferreus_bbfmm.pyi): the same class names
re-exported from the B200 mirror, so ``from ferreus_bbfmm import FmmTree, FmmKernelType, KernelParams`` (the
reference examples, py_ferreus_bbfmm/examples/*.py) runs unmodified on libferreus_b200.so.  No compute lives here."""
from ferreus_rbf_rs_b200.bbfmm import (FmmKernelType, FmmParams, FmmTree, KernelParams,  # noqa: F401
                                       M2LCompressionType, SpheroidalOrder)

__all__ = ["FmmKernelType", "FmmParams", "FmmTree", "KernelParams", "M2LCompressionType", "SpheroidalOrder"]

"""Mirror of the reference module ``ferreus_rbf.config``
(py_ferreus_rbf/src/python_bindings.rs:193-280, 570-610; ferreus_rbf/src/config.rs)."""
import enum

from .interpolant_config import RBFKernelType


class FmmCompressionType(enum.IntEnum):
    None_ = 0
    SVD = 1
    ACA = 2


class Solvers(enum.IntEnum):
    DDM = 0
    FGMRES = 1


class DDMParams:
    def __init__(self, leaf_threshold, overlap_quota, coarse_ratio, coarse_threshold):
        self.leaf_threshold = int(leaf_threshold)
        self.overlap_quota = float(overlap_quota)
        self.coarse_ratio = float(coarse_ratio)
        self.coarse_threshold = int(coarse_threshold)


class FmmParams:
    """note the leading interpolation_order (config.rs:211-253), unlike ferreus_bbfmm.FmmParams"""

    def __init__(self, interpolation_order, max_points_per_cell, compression_type, epsilon, eval_chunk_size):
        self.interpolation_order = int(interpolation_order)
        self.max_points_per_cell = int(max_points_per_cell)
        self.compression_type = FmmCompressionType(compression_type)
        self.epsilon = float(epsilon)
        self.eval_chunk_size = int(eval_chunk_size)


def _default_order(kernel_type):  # config.rs:200-207
    return {RBFKernelType.Linear: 7, RBFKernelType.ThinPlateSpline: 9, RBFKernelType.Cubic: 11}.get(
        RBFKernelType(kernel_type), 7)


class Params:
    """Params(kernel_type, *, solver_type=None, ddm_params=None, fmm_params=None, naive_solve_threshold=None,
    test_unique=None) — python_bindings.rs:570-610, defaults config.rs:141-149"""

    def __init__(self, kernel_type, *, solver_type=None, ddm_params=None, fmm_params=None,
                 naive_solve_threshold=None, test_unique=None):
        self.kernel_type = RBFKernelType(kernel_type)
        self.solver_type = Solvers.FGMRES if solver_type is None else Solvers(solver_type)
        self.ddm_params = ddm_params if ddm_params is not None else DDMParams(1024, 0.5, 0.125, 4096)
        order = _default_order(kernel_type)
        self.fmm_params = fmm_params if fmm_params is not None else \
            FmmParams(order, 256, FmmCompressionType.ACA, 10.0 ** (-order), 1024)
        self.naive_solve_threshold = 4096 if naive_solve_threshold is None else int(naive_solve_threshold)
        self.test_unique = True if test_unique is None else bool(test_unique)

"""The reference's importable module names (py_ferreus_bbfmm/src/lib.rs:15-24, py_ferreus_rbf/src/lib.rs:15-87) resolve
to the B200 mirror: same classes, submodules registered in sys.modules, import statements of the reference examples."""
import sys

import numpy as np
import pytest


def test_ferreus_bbfmm_names():
    import ferreus_bbfmm
    import ferreus_rbf_rs_b200 as fb
    from ferreus_bbfmm import FmmKernelType, FmmParams, FmmTree, KernelParams, M2LCompressionType, SpheroidalOrder
    assert FmmTree is fb.FmmTree and KernelParams is fb.KernelParams and FmmParams is fb.FmmParams
    assert [m.name for m in FmmKernelType] == ["LinearRbf", "ThinPlateSplineRbf", "CubicRbf", "SpheroidalRbf",
                                               "Laplacian", "OneOverR2", "OneOverR4"]       # python_bindings.rs:66-76
    assert [m.name for m in M2LCompressionType] == ["None_", "SVD", "ACA"]                  # :88-95
    # eq_int discriminants of the Rust enum (python_bindings.rs:79-86); the .pyi stub's 3/5/7/9 are documentation only
    assert [m.name for m in SpheroidalOrder] == ["Three", "Five", "Seven", "Nine"]
    assert set(ferreus_bbfmm.__all__) >= {"FmmTree", "KernelParams", "FmmParams"}


def test_ferreus_rbf_submodules_in_sys_modules():
    import ferreus_rbf
    import ferreus_rbf_rs_b200 as fb
    for name in ("config", "interpolant_config", "progress"):
        assert sys.modules["ferreus_rbf." + name] is getattr(fb, name)
    # the import lines of py_ferreus_rbf/examples/*.py
    from ferreus_rbf import RBFInterpolator, RBFTestFunctions  # noqa: F401
    from ferreus_rbf.config import DDMParams, FmmCompressionType, FmmParams, Params, Solvers  # noqa: F401
    from ferreus_rbf.interpolant_config import (Drift, FittingAccuracy, FittingAccuracyType,  # noqa: F401
                                                InterpolantSettings, RBFKernelType, SpheroidalOrder)
    from ferreus_rbf.progress import DuplicatesRemoved, Message, Progress, SolverIteration  # noqa: F401
    assert RBFInterpolator is fb.RBFInterpolator and ferreus_rbf.GlobalTrend is fb.GlobalTrend
    p = Params(RBFKernelType.Cubic)
    assert p.fmm_params.interpolation_order == 11 and p.naive_solve_threshold == 4096   # config.rs:141-149, 200-207
    assert [int(d) for d in Drift] == [0, 1, 2, 3]


def test_test_functions_known_values():
    from ferreus_rbf import RBFTestFunctions
    # rbf_test_functions.rs:44-99 at the origin: 3/4 e^{-2} + 3/4 e^{-1/49 - 1/10} + 1/2 e^{-58/4} - 1/5 e^{-65}
    want = 0.75 * np.exp(-2.0) + 0.75 * np.exp(-1.0 / 49.0 - 0.1) + 0.5 * np.exp(-14.5) - 0.2 * np.exp(-65.0)
    assert abs(float(RBFTestFunctions.franke_2d(np.zeros((1, 2)))[0, 0]) - want) < 1e-15
    want3 = (0.75 * np.exp(-3.0) + 0.75 * np.exp(-1.0 / 49.0 - 0.2) + 0.5 * np.exp(-(49 + 9 + 25) / 4.0)
             - 0.2 * np.exp(-(16 + 49 + 25.0)))
    assert abs(float(RBFTestFunctions.f1_3d(np.zeros((1, 3)))[0, 0]) - want3) < 1e-15


@pytest.mark.gpu
def test_reference_example_flow_runs_through_the_shims():
    """py_ferreus_bbfmm/examples/matrix_vector_product.py and py_ferreus_rbf/examples/franke_2d.py, minus the plots."""
    from ferreus_bbfmm import FmmKernelType, FmmTree, KernelParams
    from ferreus_rbf import RBFInterpolator, RBFTestFunctions
    from ferreus_rbf.interpolant_config import InterpolantSettings, RBFKernelType
    np.random.seed(42)
    src = np.random.random((10000, 3)) * 2 - 1
    w = np.random.random((10000, 2))
    tree = FmmTree(src, 7, KernelParams(FmmKernelType.LinearRbf), True, True)
    tree.set_weights(w)
    vals = tree.evaluate(w, src.copy())
    assert vals.shape == (10000, 2) and np.isfinite(vals).all()
    pts = np.random.random((100, 2))
    rbfi = RBFInterpolator(pts, RBFTestFunctions.franke_2d(pts), InterpolantSettings(RBFKernelType.ThinPlateSpline))
    fit = np.asarray(rbfi.evaluate(pts)).reshape(-1)
    assert np.abs(fit - RBFTestFunctions.franke_2d(pts)[:, 0]).max() < 0.01   # ferreus_rbf/src/lib.rs:42-89 doctest bar

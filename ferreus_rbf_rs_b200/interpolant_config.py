"""Mirror of the reference module ``ferreus_rbf.interpolant_config``
(py_ferreus_rbf/src/python_bindings.rs:399-562, ferreus_rbf/src/interpolant_config.rs)."""
import enum


class RBFKernelType(enum.IntEnum):
    Linear = 0
    ThinPlateSpline = 1
    Cubic = 2
    Spheroidal = 3


class Drift(enum.IntEnum):
    None_ = 0
    Constant = 1
    Linear = 2
    Quadratic = 3


class SpheroidalOrder(enum.IntEnum):
    Three = 3
    Five = 5
    Seven = 7
    Nine = 9


class FittingAccuracyType(enum.IntEnum):
    Relative = 0
    Absolute = 1


class FittingAccuracy:
    def __init__(self, tolerance, tolerance_type):
        self.tolerance = float(tolerance)
        self.tolerance_type = FittingAccuracyType(tolerance_type)


class InterpolantSettings:
    """InterpolantSettings(kernel_type, *, drift=None, nugget=None, spheroidal_order=None, base_range=None,
    total_sill=None, fitting_accuracy=None) — python_bindings.rs:456-517"""

    def __init__(self, kernel_type, *, drift=None, nugget=None, spheroidal_order=None, base_range=None,
                 total_sill=None, fitting_accuracy=None):
        self.kernel_type = RBFKernelType(kernel_type)
        self.drift = None if drift is None else Drift(drift)
        self.nugget = 0.0 if nugget is None else float(nugget)
        self.spheroidal_order = SpheroidalOrder.Three if spheroidal_order is None else SpheroidalOrder(spheroidal_order)
        self.base_range = 1.0 if base_range is None else float(base_range)
        self.total_sill = 1.0 if total_sill is None else float(total_sill)
        self.fitting_accuracy = fitting_accuracy if fitting_accuracy is not None else \
            FittingAccuracy(1e-6, FittingAccuracyType.Relative)

"""Oracle restatement of the reference kernel zoo (test infrastructure only).

Follows ferreus_rbf_utils/src/rbf_kernels.rs:25-317, non_rbf_kernels.rs:20-163,
constants.rs:13-50, utils.rs:230-262 (distance_sq accumulates x,y,z in order).
Registry order (utils.rs:558-571) defines the integer kernel ids used by the
C ABI.
"""
import numpy as np

EPS = np.finfo(np.float64).eps

# registry order, ferreus_rbf_utils/src/utils.rs:558-571
LINEAR, TPS, CUBIC, SPH3, SPH5, SPH7, SPH9, LAPLACIAN, ONE_OVER_R2, ONE_OVER_R4 = range(10)
KERNEL_NAMES = ["LinearRbf", "ThinPlateSplineRbf", "CubicRbf", "Spheroidal3Rbf", "Spheroidal5Rbf",
                "Spheroidal7Rbf", "Spheroidal9Rbf", "Laplacian", "OneOverR2", "OneOverR4"]

# constants.rs:21-50  (inflexion_point, linear_slope, range_scaling, inv_y_intercept)
SPHEROIDAL_CONSTANTS = {
    SPH3: (0.5000000000, 0.7500000000, 2.6798340586, 0.8734640537, 1),
    SPH5: (0.4082482905, 1.0206207262, 1.5822795750, 0.8575980168, 2),
    SPH7: (0.3535533906, 1.2374368671, 1.2008676644, 0.8494862533, 3),
    SPH9: (0.3162277660, 1.4230249471, 1.0000000000, 0.8445585690, 4),
}


class Kernel:
    """kernel_type in registry order + base_range/total_sill (kernel_helpers.rs:17-36)."""

    def __init__(self, kernel_type, base_range=1.0, total_sill=1.0):
        self.kernel_type = int(kernel_type)
        self.base_range = float(base_range)
        self.total_sill = float(total_sill)
        if self.kernel_type in SPHEROIDAL_CONSTANTS:
            ip, slope, scaling, inv_y, pw = SPHEROIDAL_CONSTANTS[self.kernel_type]
            s = scaling / self.base_range                      # rbf_kernels.rs:229-238
            self.s2 = s * s
            self.ip2 = ip * ip
            self.near_slope = self.total_sill * slope * s
            self.far_coef = self.total_sill * inv_y
            self.pow = pw

    # ---- value from squared distance -------------------------------------------------
    def eval_r2(self, r2):
        r2 = np.asarray(r2, dtype=np.float64)
        kt = self.kernel_type
        if kt == LINEAR:                                        # rbf_kernels.rs:30-40
            return -np.sqrt(r2)
        r = np.sqrt(r2)
        if kt == TPS:                                           # :77-83  r.powi(2)*r.ln(), 0 if r<eps
            with np.errstate(divide="ignore", invalid="ignore"):
                v = (r * r) * np.log(r)
            return np.where(np.abs(r) < EPS, 0.0, v)
        if kt == CUBIC:                                         # :129-133 r.powi(3)
            return r * r * r
        if kt in SPHEROIDAL_CONSTANTS:                          # :243-256
            sr2 = self.s2 * r2
            t = 1.0 + sr2
            t2 = t * t                                          # powi: binary exponentiation, as LLVM lowers it
            tp = {1: t, 2: t2, 3: t2 * t, 4: t2 * t2}[self.pow]
            far = self.far_coef / (tp * np.sqrt(t))
            near = self.total_sill - self.near_slope * r
            return np.where(sr2 <= self.ip2, near, far)
        with np.errstate(divide="ignore", invalid="ignore"):
            if kt == LAPLACIAN:                                 # non_rbf_kernels.rs:22-29
                v = 1.0 / r
            elif kt == ONE_OVER_R2:                             # :72-79
                v = 1.0 / (r * r)
            elif kt == ONE_OVER_R4:                             # :124-131
                v = 1.0 / ((r * r) * (r * r))
            else:
                raise ValueError("unknown kernel")
        return np.where(np.abs(r) < EPS, 0.0, v)

    def phi(self, r):
        r = np.asarray(r, dtype=np.float64)
        return self.eval_r2(r * r)

    # ---- value + gradient factor:  grad = factor * (target - source) --------------------
    def value_and_grad_factor(self, r2):
        """Returns (value, factor) following evaluate_value_gradient of each kernel.
        Gradient is zero when r2 <= eps (rbf_kernels.rs:49-52 etc.)."""
        r2 = np.asarray(r2, dtype=np.float64)
        kt = self.kernel_type
        small = r2 <= EPS
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.sqrt(r2)
            if kt == LINEAR:
                val = -r
                fac = -1.0 / r
            elif kt == TPS:
                val = np.where(small, 0.0, r2 * np.log(r))
                fac = 2.0 * np.log(r) + 1.0
            elif kt == CUBIC:
                val = np.where(small, 0.0, r2 * r)
                fac = 3.0 * r
            elif kt in SPHEROIDAL_CONSTANTS:
                val = self.eval_r2(r2)
                sr2 = self.s2 * r2
                t = 1.0 + sr2
                p = self.pow + 0.5
                fac_far = -2.0 * p * self.s2 * self.far_coef / np.power(t, p + 1.0)
                fac_near = -self.near_slope * (1.0 / r)
                fac = np.where(sr2 <= self.ip2, fac_near, fac_far)
            elif kt == LAPLACIAN:
                inv_r = 1.0 / r
                val = np.where(small, 0.0, inv_r)
                fac = -(inv_r * inv_r * inv_r)
            elif kt == ONE_OVER_R2:
                val = np.where(small, 0.0, 1.0 / r2)
                fac = -2.0 * (1.0 / (r2 * r2))
            elif kt == ONE_OVER_R4:
                val = np.where(small, 0.0, 1.0 / (r2 * r2))
                fac = -4.0 * (1.0 / (r2 * r2 * r2))
            else:
                raise ValueError("unknown kernel")
        fac = np.where(small, 0.0, fac)
        return val, fac

    # ---- dense helpers ----------------------------------------------------------------
    def matrix(self, targets, sources):
        """K[i, j] = k(targets_i, sources_j)  (ferreus_bbfmm/src/utils.rs:63-88)."""
        targets = np.atleast_2d(np.asarray(targets, dtype=np.float64))
        sources = np.atleast_2d(np.asarray(sources, dtype=np.float64))
        r2 = np.zeros((targets.shape[0], sources.shape[0]))
        for d in range(targets.shape[1]):                       # distance_sq order x, y, z
            diff = targets[:, d][:, None] - sources[:, d][None, :]
            r2 = r2 + diff * diff
        return self.eval_r2(r2)


def dense_matvec(kernel, targets, sources, weights, block=2048, with_gradients=False):
    """Exact O(N*M) summation in blocks: the ground truth for accuracy tests."""
    targets = np.atleast_2d(np.asarray(targets, dtype=np.float64))
    sources = np.atleast_2d(np.asarray(sources, dtype=np.float64))
    weights = np.asarray(weights, dtype=np.float64)
    if weights.ndim == 1:
        weights = weights[:, None]
    m, d = targets.shape
    nrhs = weights.shape[1]
    out = np.zeros((m, nrhs))
    grads = np.zeros((m, nrhs * d)) if with_gradients else None
    for i0 in range(0, m, block):
        t = targets[i0:i0 + block]
        diffs = [t[:, k][:, None] - sources[:, k][None, :] for k in range(d)]
        r2 = np.zeros((t.shape[0], sources.shape[0]))
        for df in diffs:
            r2 = r2 + df * df
        if with_gradients:
            val, fac = kernel.value_and_grad_factor(r2)
            out[i0:i0 + block] = val @ weights
            for k in range(d):
                g = (fac * diffs[k]) @ weights
                for j in range(nrhs):
                    grads[i0:i0 + block, j * d + k] = g[:, j]
        else:
            out[i0:i0 + block] = kernel.eval_r2(r2) @ weights
    return (out, grads) if with_gradients else out

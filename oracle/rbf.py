"""Oracle restatement of the ferreus_rbf solve path (test infrastructure only).

  evaluate_monomials        polynomials.rs:15-62           get_cheb_cube_scaling_factors  common.rs:299-320
  farthest_point_sampling   common.rs:246-287             remove_duplicates              rbf.rs:1391-1467
  Domain.factorise/solve    domain.rs:153-467             DDMTree                        domain_decomposition.rs:67-346
  schwarz_preconditioner    schwarz.rs:32-155             fgmres / schwarz_ddm_solver    iterative_solvers.rs:38-281
  fast_matrix_vector_product rbf.rs:1338-1379             RBFInterpolator fit/evaluate   rbf.rs:317-582, 676-703, 1180-1270

Dense factorizations are faer 0.23.2 in the reference (col-piv QR domain.rs:187,219; full-piv LU
polynomials.rs:120-123; partial-piv LU domain.rs:368; LLT linalg.rs:200-211; thin QR rbf.rs:493-495);
scipy/numpy (LAPACK) stand in.  Neighbour order in the overlap selection comes from rstar's traversal in
the reference (rtree.rs:76-89); it is ascending leaf index here (and in the product).
"""
import math

import numpy as np
import scipy.linalg as sla

from . import bbfmm as obb
from . import chebyshev
from .kernels import Kernel

LINEAR, TPS, CUBIC, SPHEROIDAL = 0, 1, 2, 3             # interpolant_config.rs RBFKernelType
DRIFT_NONE, DRIFT_CONSTANT, DRIFT_LINEAR, DRIFT_QUADRATIC = 0, 1, 2, 3
RELATIVE, ABSOLUTE = 0, 1
SOLVER_DDM, SOLVER_FGMRES = 0, 1


class InterpolantSettings:
    """interpolant_config.rs:179-264"""

    def __init__(self, kernel_type, drift=None, nugget=0.0, spheroidal_order=3, base_range=1.0, total_sill=1.0,
                 tolerance=1e-6, tolerance_type=RELATIVE):
        self.kernel_type = kernel_type
        min_drift = {LINEAR: DRIFT_CONSTANT, TPS: DRIFT_LINEAR, CUBIC: DRIFT_LINEAR, SPHEROIDAL: DRIFT_NONE}[kernel_type]
        self.drift = min_drift if drift is None else drift
        self.nugget = nugget
        self.spheroidal_order = spheroidal_order
        self.base_range = base_range
        self.total_sill = total_sill
        self.tolerance = tolerance
        self.tolerance_type = tolerance_type
        self.basis_size = 0
        self.polynomial_degree = -1

    def set_basis_size(self, dim):
        deg = {DRIFT_NONE: -1, DRIFT_CONSTANT: 0, DRIFT_LINEAR: 1, DRIFT_QUADRATIC: 2}[self.drift]
        min_deg = {LINEAR: 0, TPS: 1, CUBIC: 1, SPHEROIDAL: -1}[self.kernel_type]
        if deg < min_deg:
            raise ValueError(f"Min degree for kernel: {min_deg}")
        k = deg + 1
        self.basis_size = 0 if deg < 0 else {1: k, 2: k * (k + 1) // 2, 3: k * (k + 1) * (k + 2) // 6}[dim]
        self.polynomial_degree = deg

    def kernel(self):
        idx = {LINEAR: 0, TPS: 1, CUBIC: 2}.get(self.kernel_type)
        if idx is None:
            idx = {3: 3, 5: 4, 7: 5, 9: 6}[self.spheroidal_order]
        return Kernel(idx, self.base_range, self.total_sill)


class Params:
    """config.rs:98-253 defaults"""

    def __init__(self, kernel_type, solver_type=SOLVER_FGMRES, leaf_threshold=1024, overlap_quota=0.5, coarse_ratio=0.125,
                 coarse_threshold=4096, interpolation_order=None, max_points_per_cell=256,
                 compression_type=chebyshev.COMPRESSION_ACA, epsilon=None, naive_solve_threshold=4096, test_unique=True):
        self.solver_type = solver_type
        self.leaf_threshold, self.overlap_quota = leaf_threshold, overlap_quota
        self.coarse_ratio, self.coarse_threshold = coarse_ratio, coarse_threshold
        default_order = {LINEAR: 7, TPS: 9, CUBIC: 11}.get(kernel_type, 7)
        self.interpolation_order = default_order if interpolation_order is None else interpolation_order
        self.max_points_per_cell = max_points_per_cell
        self.compression_type = compression_type
        self.epsilon = 10.0 ** (-default_order) if epsilon is None else epsilon
        self.naive_solve_threshold = naive_solve_threshold
        self.test_unique = test_unique


# ------------------------------------------------------------------------------------------ helpers
def get_cheb_cube_scaling_factors(points):
    lo, hi = points.min(axis=0), points.max(axis=0)
    scale = (hi - lo) / 2.0
    scale[scale == 0.0] = 1.0
    return (hi + lo) / 2.0, scale


def evaluate_monomials(points, degree, basis_size, translation, scale):
    sp = (points - translation[None, :]) / scale[None, :]
    n, d = sp.shape
    m = np.zeros((n, basis_size))
    if basis_size == 0:
        return m
    m[:, 0] = 1.0
    if degree >= 1:
        m[:, 1:1 + d] = sp
    if degree == 2:
        k = 1 + d
        for i in range(d):
            for j in range(i, d):
                m[:, k] = sp[:, i] * sp[:, j]
                k += 1
    return m


def evaluate_monomial_gradients(points, poly, degree, translation, scale):
    n, dims = points.shape
    nrhs = poly.shape[1]
    sp = (points - translation[None, :]) / scale[None, :]
    g = np.zeros((n, nrhs * dims))
    if degree >= 1:
        for r in range(nrhs):
            for d in range(dims):
                g[:, r * dims + d] = poly[1 + d, r] / scale[d]
    if degree == 2:
        k = 1 + dims
        for i in range(dims):
            for j in range(i, dims):
                for r in range(nrhs):
                    c = poly[k, r]
                    if i == j:
                        g[:, r * dims + i] += c * (2.0 * sp[:, i] / scale[i])
                    else:
                        g[:, r * dims + i] += c * (sp[:, j] / scale[i])
                        g[:, r * dims + j] += c * (sp[:, i] / scale[j])
                k += 1
    return g


def farthest_point_sampling(points, num_wanted, seed_index):
    n = points.shape[0]
    selected = [seed_index]
    is_sel = np.zeros(n, dtype=bool)
    is_sel[seed_index] = True
    min_d = np.full(n, np.inf)
    for _ in range(1, num_wanted):
        last = selected[-1]
        diff = points - points[last][None, :]
        r2 = np.zeros(n)
        for d in range(points.shape[1]):
            r2 = r2 + diff[:, d] * diff[:, d]
        dist = np.sqrt(r2)
        upd = (~is_sel) & (dist < min_d)
        min_d[upd] = dist[upd]
        cand = np.where(is_sel, -np.inf, min_d)
        best = int(np.argmax(cand))            # first index of the maximum, as the strict `>` scan (common.rs:272-280)
        if not (cand[best] > -1.0):
            best = 0
        selected.append(best)
        is_sel[best] = True
    return selected


def duplicate_cutoff_distance(h_ref, kernel):
    """rbf.rs:1391-1416; the root is found by bisection here (roots::find_root_inverse_quadratic in the reference)."""
    eps = np.finfo(float).eps
    phi = lambda r: float(kernel.phi(r))
    phi0, phih = phi(0.0), phi(h_ref)
    target = eps * abs(phih - phi0)
    resid = lambda r: abs(phi(r) - phi0) - target
    if resid(h_ref) <= 0.0:
        return h_ref
    lo, hi = 0.0, h_ref
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if resid(mid) > 0.0:
            hi = mid
        else:
            lo = mid
        if hi - lo <= 1e-13 * hi:
            break
    return 0.5 * (lo + hi)


def remove_duplicates(points, kernel):
    """rbf.rs:1418-1467: greedy in index order, infinity-norm radius, inclusive."""
    ext = points.max(axis=0) - points.min(axis=0)
    tol = duplicate_cutoff_distance(float(np.max(np.abs(ext))), kernel)
    n = points.shape[0]
    visited = np.zeros(n, dtype=bool)
    order = np.argsort(points[:, 0], kind="stable")
    xs = points[order, 0]
    keep = []
    for i in range(n):
        if visited[i]:
            continue
        keep.append(i)
        a = np.searchsorted(xs, points[i, 0] - tol, side="left")
        b = np.searchsorted(xs, points[i, 0] + tol, side="right")
        cand = order[a:b]
        near = cand[np.all(np.abs(points[cand] - points[i][None, :]) <= tol, axis=1)]
        visited[near] = True
    return np.array(keep, dtype=np.int64)


# ------------------------------------------------------------------------------------------ domain
class GlobalTrend:
    """global_trend.rs:36-126: One{major_ratio} | Two{rotation_angle, major_ratio, minor_ratio} |
    Three{dip, dip_direction, pitch, major_ratio, semi_major_ratio, minor_ratio}; angles in degrees."""

    def __init__(self, dim, angles, ratios):
        self.dim, self.angles, self.ratios = dim, list(angles), list(ratios)

    @staticmethod
    def one(major_ratio):
        return GlobalTrend(1, [], [major_ratio])

    @staticmethod
    def two(rotation_angle, major_ratio, minor_ratio):
        return GlobalTrend(2, [rotation_angle], [major_ratio, minor_ratio])

    @staticmethod
    def three(dip, dip_direction, pitch, major_ratio, semi_major_ratio, minor_ratio):
        return GlobalTrend(3, [dip, dip_direction, pitch], [major_ratio, semi_major_ratio, minor_ratio])


def _rot_z(a, d):
    m = np.eye(d + 1)
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
    return m


class GlobalTrendTransform:
    """global_trend.rs:128-287: affine = (translate_back * scale * rotation * translate)^T applied to homogeneous
    ROW vectors; inverse by LU."""

    def __init__(self, center, trend):
        d = trend.dim
        t = np.eye(d + 1)
        tb = np.eye(d + 1)
        t[:d, d] = -np.asarray(center)
        tb[:d, d] = np.asarray(center)
        scale = np.eye(d + 1)
        for i in range(d):
            scale[i, i] = 1.0 / trend.ratios[i]
        if d == 1:
            rot = np.eye(2)
        elif d == 2:
            rot = _rot_z(-np.radians(trend.angles[0]), 2)
        else:
            dipr, dipdirr, pitchr = (-np.radians(a) for a in trend.angles)
            rot_z = _rot_z(dipdirr, 3)
            rot_x = np.eye(4)
            rot_x[1, 1], rot_x[1, 2], rot_x[2, 1], rot_x[2, 2] = np.cos(dipr), np.sin(dipr), -np.sin(dipr), np.cos(dipr)
            rot_z2 = _rot_z(pitchr, 3)
            rot = rot_z2 @ rot_x @ rot_z
        self.affine = (tb @ scale @ rot @ t).T.copy()
        self.inverse = np.linalg.inv(self.affine)
        self.dim = d

    def transform_points(self, pts):
        h = np.hstack([pts, np.ones((pts.shape[0], 1))])
        return np.ascontiguousarray((h @ self.affine)[:, :self.dim])

    def inverse_transform_points(self, pts):
        h = np.hstack([pts, np.ones((pts.shape[0], 1))])
        return np.ascontiguousarray((h @ self.inverse)[:, :self.dim])

    def linear_part(self):
        return self.affine[:self.dim, :self.dim].copy()


class Domain:
    def __init__(self, indices):
        self.idx = np.array(indices, dtype=np.int64)
        self.mask = np.zeros(0, dtype=bool)
        self.extents = None
        self.solve_for_poly = False
        self.q_top = None
        self.rank = 0

    def factorise(self, points, settings, solve_for_poly, gt=None):
        kern = settings.kernel()
        dp = points[self.idx]
        n = len(self.idx)
        if settings.basis_size != 0:
            tr, sc = get_cheb_cube_scaling_factors(dp)
            mp = gt.inverse_transform_points(dp) if gt is not None else dp       # domain.rs:169-175
            mono = evaluate_monomials(mp, settings.polynomial_degree, settings.basis_size, tr, sc)
            _, r, piv = sla.qr(mono, mode="economic", pivoting=True)          # domain.rs:187
            diag = np.abs(np.diag(r))
            rank = int(np.sum(diag > 1e-10 * diag[0]))
            cols = np.sort(piv[:rank])
            full = mono[:, cols]
            _, _, piv2 = sla.qr(full.T, mode="economic", pivoting=True)       # domain.rs:219
            special = np.sort(piv2[:rank])
            sset = set(special.tolist())
            non_special = np.array([i for i in range(n) if i not in sset], dtype=np.int64)
            order = np.concatenate([special, non_special])
            self.idx = self.idx[order]
            self.mask = self.mask[:n][order] if len(self.mask) >= n else self.mask
            sp_mono, ns_mono = full[special], full[non_special]
            sdp = points[self.idx]
            a = kern.matrix(sdp, sdp) + settings.nugget * np.eye(n)
            lag = np.linalg.solve(sp_mono, np.eye(rank))                      # polynomials.rs:118-124
            self.q_top = -(ns_mono @ lag).T                                   # rank x m
            q = self.q_top
            a11, a12, a21, a22 = a[:rank, :rank], a[:rank, rank:], a[rank:, :rank], a[rank:, rank:]
            lhs = q.T @ (a11 @ q) + q.T @ a12 + a21 @ q + a22
            self.rank = rank
            if solve_for_poly:
                self.solve_for_poly = True
                self.a_special_rows = a[:rank].copy()
                self.sp_mono = sp_mono
        else:
            lhs = kern.matrix(dp, dp) + settings.nugget * np.eye(n)
        try:
            self.chol = sla.cho_factor(lhs, lower=True)
            self.lu = None
        except np.linalg.LinAlgError:                                          # LBLT fallback, domain.rs:63-68
            self.chol = None
            self.lu = sla.lu_factor(lhs)

    def solve(self, values):
        d = values[self.idx]
        if self.q_top is not None:
            rhs = self.q_top.T @ d[:self.rank] + d[self.rank:]
        else:
            rhs = d
        gamma = sla.cho_solve(self.chol, rhs) if self.chol is not None else sla.lu_solve(self.lu, rhs)
        if self.q_top is not None:
            coeff = np.concatenate([self.q_top @ gamma, gamma], axis=0)
        else:
            coeff = gamma
        poly = None
        if self.solve_for_poly:
            r = d[:self.rank] - self.a_special_rows @ coeff
            poly = np.linalg.solve(self.sp_mono, r)
        return coeff, poly


class Level:
    def __init__(self, point_indices):
        self.point_indices = np.array(point_indices, dtype=np.int64)
        self.leaf_domains = []


def _argmax_first_positive(v):
    best, best_val = 0, 0.0
    for i, x in enumerate(v):
        if x > best_val:
            best_val, best = x, i
    return best


class DDMTree:
    def __init__(self, points, settings, leaf_threshold, overlap_quota, coarse_ratio, coarse_threshold, factorise=True,
                 gt=None):
        n, dim = points.shape
        self.levels = []
        active = np.arange(n, dtype=np.int64)
        while len(active) > coarse_threshold:
            root = Domain(active)
            root.extents = np.concatenate([points[active].min(axis=0), points[active].max(axis=0)])
            queue = [root]
            level = Level(active)
            coarse_pts = []
            qi = 0
            while qi < len(queue):
                cur = queue[qi]
                qi += 1
                idx = cur.idx
                nd = len(idx)
                cp = points[idx]
                axis = _argmax_first_positive(cp.max(axis=0) - cp.min(axis=0))
                order = np.argsort(cp[:, axis], kind="stable")
                sorted_idx = idx[order]
                mid = nd // 2
                left, right = Domain(np.sort(sorted_idx[:mid])), Domain(np.sort(sorted_idx[mid:]))
                mid_coord = points[sorted_idx[mid], axis]
                left.extents = cur.extents.copy()
                left.extents[axis + dim] = mid_coord
                right.extents = cur.extents.copy()
                right.extents[axis] = mid_coord
                if nd + nd * overlap_quota >= 2.0 * leaf_threshold:
                    queue.extend([left, right])
                else:
                    for dmn in (left, right):
                        dmn.mask = np.ones(len(dmn.idx), dtype=bool)
                    level.leaf_domains.extend([left, right])
            leaves = level.leaf_domains
            num_coarse = int(math.ceil(math.ceil(len(active) * coarse_ratio) / len(leaves)))
            internal = [d.idx.copy() for d in leaves]
            for i, dmn in enumerate(leaves):
                ii = internal[i]
                ip = points[ii]
                sample = min(len(ii), num_coarse)
                centroid = np.array([ip[:, c].sum() / len(ii) for c in range(dim)])
                diff = ip - centroid[None, :]
                r2 = np.zeros(len(ii))
                for c in range(dim):
                    r2 = r2 + diff[:, c] * diff[:, c]
                centre_index = int(np.argmin(np.sqrt(r2)))
                sel = farthest_point_sampling(ip, sample, centre_index)
                coarse_pts.extend(sorted(ii[s] for s in sel))
                # neighbours: boxes that intersect (touching counts), excluding self — rtree.rs:76-89
                lo, hi = dmn.extents[:dim], dmn.extents[dim:]
                nb = [j for j, o in enumerate(leaves) if j != i
                      and np.all(o.extents[:dim] <= hi) and np.all(o.extents[dim:] >= lo)]
                num_overlap = int(math.ceil(len(dmn.idx) * 2 * overlap_quota))
                nidx = np.concatenate([internal[j] for j in nb]) if nb else np.zeros(0, dtype=np.int64)
                pts = points[nidx]
                clipped = np.maximum(np.minimum(pts, hi[None, :]), lo[None, :])
                df = pts - clipped
                r2 = np.zeros(len(nidx))
                for c in range(dim):
                    r2 = r2 + df[:, c] * df[:, c]
                order = np.argsort(np.sqrt(r2), kind="stable")[:min(num_overlap, len(nidx))]
                dmn.idx = np.concatenate([dmn.idx, nidx[order]])
                dmn.mask = np.concatenate([dmn.mask, np.zeros(len(order), dtype=bool)])
            if factorise:
                for dmn in leaves:
                    dmn.factorise(points, settings, False, gt)
            self.levels.append(level)
            active = np.array(sorted(coarse_pts), dtype=np.int64)
        coarse = Level(active)
        cd = Domain(active)
        cd.mask = np.ones(len(active), dtype=bool)
        if factorise:
            cd.factorise(points, settings, settings.basis_size != 0, gt)
        coarse.leaf_domains.append(cd)
        self.levels.append(coarse)


# ------------------------------------------------------------------------------------------ solver
def givens_rotation(f, g):
    """iterative_solvers.rs:192-232"""
    if g == 0.0:
        return 1.0, 0.0, f
    if f == 0.0:
        return 0.0, math.copysign(1.0, g), abs(g)
    r = math.copysign(math.sqrt(f * f + g * g), f)
    return abs(f) / abs(r), g / r, r


def fgmres(matvec, b, precon, max_outer, max_inner, tolerance, tolerance_type, callback=None):
    """iterative_solvers.rs:38-173 (one right-hand side, column vector b)."""
    n = b.shape[0]
    x = np.zeros(n)
    r = b - matvec(x)
    beta = np.max(np.abs(r)) if tolerance_type == ABSOLUTE else np.linalg.norm(r)
    iteration = 1
    for _outer in range(max_outer):
        v = np.zeros((n, max_inner + 1))
        h = np.zeros((max_inner + 1, max_inner))
        z = np.zeros((n, max_inner))
        g = np.zeros(max_inner + 1)
        cs, sn = np.zeros(max_inner), np.zeros(max_inner)
        r_norm = np.linalg.norm(r)
        v[:, 0] = r / r_norm
        g[0] = r_norm
        for j in range(max_inner):
            w = precon(v[:, j]) if precon is not None else v[:, j].copy()
            z[:, j] = w
            wj = matvec(w)
            for i in range(j + 1):
                hij = float(np.dot(v[:, i], wj))
                h[i, j] = hij
                wj = wj - v[:, i] * hij
            norm = np.linalg.norm(wj)
            h[j + 1, j] = norm
            for i in range(j):
                temp = cs[i] * h[i, j] + sn[i] * h[i + 1, j]
                h[i + 1, j] = -sn[i] * h[i, j] + cs[i] * h[i + 1, j]
                h[i, j] = temp
            c, s, _ = givens_rotation(h[j, j], h[j + 1, j])
            h[j, j] = c * h[j, j] + s * h[j + 1, j]
            h[j + 1, j] = 0.0
            temp = c * g[j] + s * g[j + 1]
            g[j + 1] = -s * g[j] + c * g[j + 1]
            g[j] = temp
            cs[j], sn[j] = c, s
            if norm != 0.0:
                v[:, j + 1] = wj / norm
            res = abs(g[j + 1]) if tolerance_type == ABSOLUTE else abs(g[j + 1]) / beta
            if callback is not None:
                callback(iteration, res)
            if res < tolerance:
                y = sla.solve_triangular(h[:j + 1, :j + 1], g[:j + 1])
                return x + z[:, :j + 1] @ y
            iteration += 1
        y = sla.solve_triangular(h[:max_inner, :max_inner], g[:max_inner])
        x = x + z @ y
        r = b - matvec(x)
        res = np.max(np.abs(r)) if tolerance_type == ABSOLUTE else np.linalg.norm(r) / beta
        if res < tolerance:
            break
    return x


def schwarz_ddm_solver(matvec, rhs, precon, max_iterations, tolerance, tolerance_type, callback=None):
    """iterative_solvers.rs:234-281"""
    rg = rhs.copy()
    sg = np.zeros_like(rhs)
    beta = np.max(np.abs(rg)) if tolerance_type == ABSOLUTE else np.linalg.norm(rg)
    res = beta
    it = 0
    while res > tolerance and it < max_iterations:
        sg = sg + precon(rg)
        rg = rhs - matvec(sg)
        res = np.max(np.abs(rg)) if tolerance_type == ABSOLUTE else np.linalg.norm(rg) / beta
        it += 1
        if callback is not None:
            callback(it, res)
    return sg


class RBFInterpolator:
    """rbf.rs:317-582 (fit) and 676-703 / 1180-1270 (evaluate), global trend included (rbf.rs:361-371, 477-484,
    579-581, 599-615, 1183-1229)."""

    def __init__(self, points, values, settings, params=None, callback=None, dense_matvec=False, global_trend=None):
        points = np.array(points, dtype=np.float64)
        values = np.array(values, dtype=np.float64)
        if values.ndim == 1:
            values = values[:, None]
        dim = points.shape[1]
        settings.set_basis_size(dim)
        self.settings = settings
        self.params = params if params is not None else Params(settings.kernel_type)
        kern = settings.kernel()
        self.kernel = kern
        if self.params.test_unique:
            keep = remove_duplicates(points, kern)
            self.num_duplicates = points.shape[0] - len(keep)
            if len(keep) != points.shape[0]:
                points, values = points[keep], values[keep]
        self.gt = None
        if global_trend is not None:                         # rbf.rs:361-371: centre = mean of the unique points
            self.gt = GlobalTrendTransform(points.mean(axis=0), global_trend)
            points = self.gt.transform_points(points)
        gt = self.gt
        self.points, self.values = points, values
        n, m = points.shape[0], settings.basis_size
        self.translation, self.scale = (get_cheb_cube_scaling_factors(points) if m else (None, None))
        self.iterations = 0
        if n < self.params.naive_solve_threshold:
            dom = Domain(np.arange(n))
            dom.mask = np.ones(n, dtype=bool)
            dom.factorise(points, settings, True, gt)
            coeff, poly = dom.solve(values)
            pc = np.zeros_like(coeff)
            pc[dom.idx] = coeff
            self.point_coefficients, self.poly_coefficients = pc, poly
            if gt is not None:
                self.points = gt.inverse_transform_points(self.points)   # rbf.rs:579-581
            return
        p = self.params
        fmm_params = obb.FmmParams(p.max_points_per_cell, p.compression_type, p.epsilon, 1024)
        self.tree = obb.FmmTree(points, p.interpolation_order, kern, True, True, None, fmm_params)
        self.fast = None
        if dense_matvec is False:
            from . import fast
            self.fast = fast.FastFmm(self.tree)
        mono_points = gt.inverse_transform_points(points) if gt is not None else points   # rbf.rs:477-484
        P = evaluate_monomials(mono_points, settings.polynomial_degree, m, self.translation, self.scale) if m else None
        Qp = np.linalg.qr(P)[0] if m else None
        self.ddm = DDMTree(points, settings, p.leaf_threshold, p.overlap_quota, p.coarse_ratio, p.coarse_threshold,
                           gt=gt)
        nugget = settings.nugget

        def matvec_partial(w, idx=None):                     # rbf.rs:1338-1379
            res = np.zeros(n + m)
            ev = np.arange(n) if idx is None else idx
            fmm = self.fast.matvec(w[:n])[:, 0] if self.fast is not None else \
                self.kernel.matrix(points, points) @ w[:n]
            res[ev] = fmm[ev] + nugget * w[ev]
            if m:
                res[ev] += P[ev] @ w[n:]
            return res

        def precon(rg):                                       # schwarz.rs:32-79
            sl = np.zeros(n + m)
            levels = self.ddm.levels
            coarse_idx = len(levels) - 1
            cidx = levels[coarse_idx].point_indices

            def coarse(res, add_poly):
                sc_ = np.zeros(n + m)
                dom = levels[coarse_idx].leaf_domains[0]
                coeff, poly = dom.solve(res[:, None])
                sc_[dom.idx] = coeff[:, 0]
                if dom.solve_for_poly and add_poly:
                    k = poly.shape[0]
                    sc_[n + m - k:] = poly[:, 0]
                return sc_

            if coarse_idx > 0:
                for i in range(coarse_idx):
                    res = rg - matvec_partial(sl, levels[i].point_indices)
                    s1 = np.zeros(n + m)
                    for dom in levels[i].leaf_domains:
                        coeff, _ = dom.solve(res[:, None])
                        k = min(len(dom.mask), len(dom.idx))
                        sel = dom.mask[:k]
                        s1[dom.idx[:k][sel]] = coeff[:k, 0][sel]
                    if m:
                        s1[:n] -= Qp @ (Qp.T @ s1[:n])
                    sl = sl + s1
                    sl = sl + coarse(rg - matvec_partial(sl, cidx), i == coarse_idx - 1)
            else:
                sl = sl + coarse(rg - matvec_partial(sl, cidx), True)
            return sl

        self.matvec, self.precon = matvec_partial, precon
        pc = np.zeros((n, values.shape[1]))
        poly = np.zeros((m, values.shape[1])) if m else None
        self.residuals = []

        def cb(it, res):
            self.iterations = it
            self.residuals.append(res)
            if callback is not None:
                callback(it, res)

        for col in range(values.shape[1]):
            rhs = np.concatenate([values[:, col], np.zeros(m)])
            if p.solver_type == SOLVER_FGMRES:
                sol = fgmres(lambda x: matvec_partial(x), rhs, precon, 20, 5, settings.tolerance,
                             settings.tolerance_type, cb)
            else:
                sol = schwarz_ddm_solver(lambda x: matvec_partial(x), rhs, precon, 100, settings.tolerance,
                                         settings.tolerance_type, cb)
            pc[:, col] = sol[:n]
            if m:
                poly[:, col] = sol[n:]
        self.point_coefficients, self.poly_coefficients = pc, poly
        if gt is not None:
            self.points = gt.inverse_transform_points(self.points)       # rbf.rs:579-581

    def evaluate(self, targets, with_gradients=False):
        """rbf.rs:676-703 + 1180-1270: non-sparse adaptive tree on the union extents."""
        targets = np.array(targets, dtype=np.float64)
        p = self.params
        lo = np.minimum(self.points.min(axis=0), targets.min(axis=0))
        hi = np.maximum(self.points.max(axis=0), targets.max(axis=0))
        src, tgt = self.points, targets
        if self.gt is not None:                               # rbf.rs:599-615: transform points and the box corners
            d = src.shape[1]
            corners = np.array([[lo[j] if ((i >> j) & 1) == 0 else hi[j] for j in range(d)] for i in range(1 << d)])
            tc = self.gt.transform_points(corners)
            lo, hi = tc.min(axis=0), tc.max(axis=0)
            src, tgt = self.gt.transform_points(src), self.gt.transform_points(targets)
        tree = obb.FmmTree(src, p.interpolation_order, self.kernel, True, False, list(lo) + list(hi),
                           obb.FmmParams(p.max_points_per_cell, p.compression_type, p.epsilon, 1024))
        tree.set_weights(self.point_coefficients)
        res = tree.evaluate(self.point_coefficients, tgt, with_gradients)
        vals, grads = res if with_gradients else (res, None)
        if self.gt is not None and with_gradients:            # rbf.rs:1272-1298: grad_x = grad_x' B^T per rhs
            d = src.shape[1]
            bt = self.gt.linear_part().T
            g = grads.reshape(grads.shape[0], -1, d)
            grads = np.einsum('irj,jk->irk', g, bt).reshape(grads.shape)
        s = self.settings
        if s.basis_size:
            mono = evaluate_monomials(targets, s.polynomial_degree, s.basis_size, self.translation, self.scale)
            vals = vals + mono @ self.poly_coefficients
            if with_gradients:
                grads = grads + evaluate_monomial_gradients(targets, self.poly_coefficients, s.polynomial_degree,
                                                            self.translation, self.scale)
        return (vals, grads) if with_gradients else vals

    def evaluate_dense(self, targets):
        """exact evaluation of the fitted interpolant (ground truth for accuracy checks)"""
        targets = np.array(targets, dtype=np.float64)
        if self.gt is not None:
            vals = self.kernel.matrix(self.gt.transform_points(targets),
                                      self.gt.transform_points(self.points)) @ self.point_coefficients
        else:
            vals = self.kernel.matrix(targets, self.points) @ self.point_coefficients
        s = self.settings
        if s.basis_size:
            vals = vals + evaluate_monomials(targets, s.polynomial_degree, s.basis_size, self.translation,
                                             self.scale) @ self.poly_coefficients
        return vals

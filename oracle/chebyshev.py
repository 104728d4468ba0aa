"""Oracle restatement of ferreus_bbfmm/src/chebyshev.rs and aca.rs (test infrastructure only).

Dense factorizations (thin QR, SVD) are the un-vendored crate faer 0.23.2 in the reference
(aca.rs:175-189, chebyshev.rs:763); numpy.linalg (LAPACK) stands in for it here.  The
truncated products U*Vt are unique up to round-off whenever the cut-off rank agrees.
"""
import itertools
import math

import numpy as np

from .kernels import Kernel

COMPRESSION_NONE, COMPRESSION_SVD, COMPRESSION_ACA = 0, 1, 2


def cartesian_product(values, ncols):
    """utils.rs:122-134: column 0 slowest."""
    return np.array(list(itertools.product(values, repeat=ncols)))


def argsort_stable(data):
    return sorted(range(len(data)), key=lambda i: data[i])


def generate_chebyshev_nodes(p):
    """chebyshev.rs:32-40 (ascending)."""
    return np.array([math.cos(math.pi * (i + 0.5) / p) for i in reversed(range(p))])


def evaluate_chebyshev_polynomials(ncols, x, with_derivatives=False):
    """chebyshev.rs:47-110: T[m, k] = T_k(x_m), dT likewise."""
    x = np.asarray(x, dtype=np.float64)
    T = np.ones((x.shape[0], ncols))
    dT = np.zeros((x.shape[0], ncols)) if with_derivatives else None
    if ncols > 1:
        T[:, 1] = x
        if with_derivatives:
            dT[:, 1] = 1.0
    for j in range(2, ncols):
        T[:, j] = 2.0 * x * T[:, j - 1] - T[:, j - 2]
        if with_derivatives:
            dT[:, j] = 2.0 * T[:, j - 1] + 2.0 * x * dT[:, j - 1] - dT[:, j - 2]
    return T, dT


def calculate_sn(tn_x, polynomial_nodes, p):
    """chebyshev.rs:114-127"""
    return ((tn_x @ polynomial_nodes.T) * 2.0 - 1.0) / float(p)


def calculate_dsn_dx(dtn_x, polynomial_nodes, p):
    """chebyshev.rs:130-142"""
    return (dtn_x @ polynomial_nodes.T) * (2.0 / float(p))


def get_m2m_transfer_matrices(p, nodes, polynomial_nodes, dim):
    """chebyshev.rs:146-241: M2M[c] = (kron over axes, axis 0 slowest)^T, bit j of c = axis j half."""
    child_nodes = np.array([(nodes[i] - 1.0) * 0.5 if i < p else (nodes[i - p] + 1.0) * 0.5
                            for i in range(2 * p)])
    T, _ = evaluate_chebyshev_polynomials(p, child_nodes)
    sn = calculate_sn(T, polynomial_nodes, p)
    halves = (sn[:p], sn[p:])
    out = []
    for c in range(1 << dim):
        acc = None
        for j in range(dim):
            m = halves[1 if (c >> j) & 1 else 0]
            acc = m.copy() if acc is None else np.kron(acc, m)
        out.append(np.ascontiguousarray(acc.T))
    return out


def get_m2l_vectors(dim):
    """chebyshev.rs:267-297"""
    all_vecs = cartesian_product(range(-3, 4), dim)
    ref = [row for row in cartesian_product(range(0, 4), dim)
           if row[0] >= 2 and all(row[i] <= row[i - 1] for i in range(1, dim))]
    return all_vecs, np.array(ref)


def get_permutation_lookups(dim, p, all_vecs, ref_vecs):
    """chebyshev.rs:486-585"""
    axis_orders = [list(o) for o in itertools.permutations(range(dim))]
    axis_signs = cartesian_product([-1, 1], dim)
    multi = cartesian_product(range(1, p + 1), dim)
    npts = multi.shape[0]
    powers = [p ** (dim - 1 - i) for i in range(dim)]
    to_k = lambda alpha: sum((alpha[i] - 1) * powers[i] for i in range(dim))

    def perm_indices(transform):
        out = [0] * npts
        for j in range(npts):
            out[to_k(transform(multi[j]))] = j
        return out

    diag = [perm_indices(lambda a, o=o: [a[i] for i in o]) for o in axis_orders]
    axial = [perm_indices(lambda a, s=s: [p - (a[i] - 1) if s[i] < 0 else a[i] for i in range(dim)])
             for s in axis_signs]
    combos = [(a, b) for a in range(len(axis_signs)) for b in range(len(axis_orders))]
    combined = [[axial[a][i] for i in diag[b]] for a, b in combos]
    inverse = [argsort_stable(c) for c in combined]

    sign_list = [list(s) for s in axis_signs]
    perm_lookups, ref_lookups = [], []
    sorted_refs = [sorted(r) for r in ref_vecs.tolist()]
    for vec in all_vecs.tolist():
        a = sign_list.index([-1 if x < 0 else 1 for x in vec])
        b = axis_orders.index(argsort_stable([-abs(x) for x in vec]))
        perm_lookups.append(combos.index((a, b)))
        sv = sorted(abs(x) for x in vec)
        ref_lookups.append(sorted_refs.index(sv) if sv in sorted_refs else 0)
    return combined, inverse, perm_lookups, ref_lookups


# --------------------------------------------------------------------------- ACA (aca.rs)
def _argmax_masked(data, mask):
    """aca.rs:146-161 (first index on ties, 0 if all zero)."""
    best, best_val = 0, 0.0
    w = np.abs(data) * mask
    for idx in range(len(w)):
        if w[idx] > best_val:
            best_val = w[idx]
            best = idx
    return best


def _seq_dot(a, b):
    """strictly sequential dot product (the summation order of the product's host code; faer's own
    order is unknowable, and the ACA pivot / stopping decisions are sensitive to it at round-off level)"""
    return float(np.cumsum(a * b)[-1])


def aca_partial_pivoting(nrows, ncols, row_fn, col_fn, epsilon):
    """aca.rs:23-136"""
    unused_rows = np.ones(nrows)
    unused_cols = np.ones(ncols)
    max_it = min(nrows, ncols)
    tol = epsilon ** 2
    u = np.zeros((nrows, max_it))
    v = np.zeros((ncols, max_it))
    residual_norm = 0.0
    i = 0
    sum_k = 0.0
    k = 0
    for _ in range(max_it):
        row = row_fn(i).copy()
        unused_rows[i] = 0
        for l in range(k):          # sequential rank-1 downdates (summation order of the product's host code)
            row -= u[i, l] * v[:, l]
        j = _argmax_masked(row, unused_cols)
        with np.errstate(divide="ignore", invalid="ignore"):
            row = row * (1.0 / row[j])
        col = col_fn(j).copy()
        unused_cols[j] = 0
        for l in range(k):
            col -= v[j, l] * u[:, l]
        i = _argmax_masked(col, unused_rows)
        if k > 0:
            sum_k = 0.0
            for l in range(k):
                sum_k += _seq_dot(u[:, l], col) * _seq_dot(v[:, l], row)
        norm_u_v_2 = _seq_dot(col, col) * _seq_dot(row, row)
        residual_norm += norm_u_v_2 + 2.0 * sum_k
        u[:, k] = col
        v[:, k] = row
        k += 1
        if norm_u_v_2 <= tol * residual_norm:
            break
    return u[:, :k].copy(), v[:, :k].copy()


def calculate_singular_values_cutoff(sigma, epsilon):
    """aca.rs:210-247: first index whose tail sum of squares < eps^2 * total."""
    sigma = np.asarray(sigma, dtype=np.float64)
    acc = 0.0
    cum = np.zeros(len(sigma))
    for idx in range(len(sigma) - 1, -1, -1):
        acc += sigma[idx] * sigma[idx]
        cum[idx] = acc
    eps_qr = cum[0] * epsilon * epsilon
    for idx in range(len(cum)):
        if cum[idx] < eps_qr:
            return idx
    return len(cum)


def recompress_aca(u_aca, v_aca, epsilon):
    """aca.rs:173-200"""
    qu, ru = np.linalg.qr(u_aca)
    qv, rv = np.linalg.qr(v_aca)
    ur, sr, vrt = np.linalg.svd(ru @ rv.T)
    rank = calculate_singular_values_cutoff(sr, epsilon)
    u = qu @ (ur[:, :rank] * sr[:rank][None, :])
    vt = vrt[:rank] @ qv.T
    return u, vt


class PrecomputeOperators:
    """bbfmm.rs:153-184 / chebyshev.rs:650-814"""

    def __init__(self, p, dim, radius, depth, kernel: Kernel, compression, epsilon):
        self.p, self.dim = p, dim
        self.num_nodes_nd = p ** dim
        self.nodes = generate_chebyshev_nodes(p)
        self.nodes_nd = cartesian_product(self.nodes, dim)
        self.polynomial_nodes, _ = evaluate_chebyshev_polynomials(p, self.nodes)
        self.m2m = get_m2m_transfer_matrices(p, self.nodes, self.polynomial_nodes, dim)
        self.all_vecs, self.ref_vecs = get_m2l_vectors(dim)
        (self.permutation_indices, self.inverse_permutations, self.permutation_lookups,
         self.reference_vector_lookups) = get_permutation_lookups(dim, p, self.all_vecs, self.ref_vecs)
        self.u, self.vt = {}, {}
        self.compression = compression
        for level in range(2, depth + 1):
            cell_length = radius / float(2 ** (level - 1))              # chebyshev.rs:702
            target_points = self.nodes_nd * (0.5 * cell_length)          # :588-600
            self.u[level], self.vt[level] = {}, {}
            for i, ref in enumerate(self.ref_vecs):
                source_points = (ref[None, :].astype(np.float64) + self.nodes_nd * 0.5) * cell_length  # :604-627
                if compression == COMPRESSION_ACA:
                    row_fn = lambda r: kernel.matrix(source_points[r:r + 1], target_points)[0]
                    col_fn = lambda c: kernel.matrix(source_points, target_points[c:c + 1])[:, 0]
                    ua, va = aca_partial_pivoting(source_points.shape[0], target_points.shape[0],
                                                  row_fn, col_fn, epsilon)
                    self.u[level][i], self.vt[level][i] = recompress_aca(ua, va, epsilon)
                elif compression == COMPRESSION_SVD:
                    a = kernel.matrix(source_points, target_points)
                    ur, sr, vrt = np.linalg.svd(a)
                    rank = calculate_singular_values_cutoff(sr, epsilon)
                    self.u[level][i] = ur[:, :rank].copy()
                    self.vt[level][i] = sr[:rank][:, None] * vrt[:rank]
                else:
                    self.u[level][i] = kernel.matrix(source_points, target_points)

    def rank(self, level, ref):
        return self.u[level][ref].shape[1]


def get_approximation_coefficients(p, points, center, length, polynomial_nodes, dim, gradients=False):
    """chebyshev.rs:831-927: rows = points, columns = p^d tensor weights (axis 0 slowest)."""
    pts = (np.asarray(points, dtype=np.float64) - np.asarray(center)[None, :]) / (length * 0.5)
    sn, dsn = [], []
    for d in range(dim):
        T, dT = evaluate_chebyshev_polynomials(p, pts[:, d], gradients)
        sn.append(calculate_sn(T, polynomial_nodes, p))
        if gradients:
            dsn.append(calculate_dsn_dx(dT, polynomial_nodes, p) * (2.0 / length))
    n = pts.shape[0]

    def tensor(mats):
        acc = mats[0]
        for m in mats[1:]:
            acc = (acc[:, :, None] * m[:, None, :]).reshape(n, -1)
        return acc

    values = tensor(sn)
    grads = None
    if gradients:
        grads = np.concatenate([tensor([dsn[d] if d == g else sn[d] for d in range(dim)])
                                for g in range(dim)], axis=1)
    return values, grads


def scale_cheb_nodes_to_cell(nodes_nd, center, length):
    """chebyshev.rs:951-968"""
    return np.asarray(center)[None, :] + (length * 0.5) * nodes_nd

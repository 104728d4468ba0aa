"""CPU tests of the library's host-side operator precompute (no GPU needed): the Chebyshev tables,
symmetry permutations, ACA + recompression and truncated SVD of host_ops.cpp against the oracle
restatement of chebyshev.rs / aca.rs."""
import ctypes as C

import numpy as np
import pytest

from ferreus_rbf_rs_b200 import _lib
from oracle import chebyshev as oc
from oracle import kernels as ok


def host_ops(p, dim, radius, depth, kt, comp, eps, base_range=1.0, total_sill=1.0):
    L = _lib.lib()
    h = C.c_void_p()
    kp = _lib.FbKernelParams(kt, base_range, total_sill)
    assert L.fb_ops_new(p, dim, radius, depth, C.byref(kp), comp, eps, C.byref(h)) == 0
    return h


def get_op(h, lvl, r, P, compressed):
    L = _lib.lib()
    rk = L.fb_ops_rank(h, lvl, r)
    u = np.zeros(P * rk)
    vt = np.zeros(rk * P)
    assert L.fb_ops_get(h, lvl, r, _lib.dptr(u), _lib.dptr(vt)) == 0
    U = u.reshape(rk, P).T
    return (U, vt.reshape(P, rk).T) if compressed else (U, None)


@pytest.mark.parametrize("p,dim", [(4, 1), (5, 2), (4, 3), (7, 3)])
def test_tables_match_oracle(p, dim):
    L = _lib.lib()
    h = host_ops(p, dim, 0.501, 1, 0, 0, 1e-3)
    oo = oc.PrecomputeOperators(p, dim, 0.501, 1, ok.Kernel(0), 0, 1e-3)
    P = p ** dim
    n_perm, n_ref = C.c_int32(), C.c_int32()
    L.fb_ops_tables(h, C.byref(n_perm), C.byref(n_ref), None, None, None, None, None)
    assert n_perm.value == len(oo.permutation_indices)
    assert n_ref.value == len(oo.ref_vecs)
    perm = np.zeros(n_perm.value * P, dtype=np.int32)
    inv = np.zeros_like(perm)
    pl = np.zeros(7 ** dim, dtype=np.int32)
    rl = np.zeros(7 ** dim, dtype=np.int32)
    cs = np.zeros(2 * p * p)
    i32 = C.POINTER(C.c_int32)
    L.fb_ops_tables(h, None, None, perm.ctypes.data_as(i32), inv.ctypes.data_as(i32), pl.ctypes.data_as(i32),
                    rl.ctypes.data_as(i32), _lib.dptr(cs))
    assert np.array_equal(perm.reshape(-1, P), np.array(oo.permutation_indices))
    assert np.array_equal(inv.reshape(-1, P), np.array(oo.inverse_permutations))
    assert np.array_equal(pl, np.array(oo.permutation_lookups))
    assert np.array_equal(rl, np.array(oo.reference_vector_lookups))
    # M2M[c] = (kron_j S_half(bit_j))^T  (chebyshev.rs:196-241)
    halves = cs.reshape(2, p, p)
    for c in range(1 << dim):
        acc = None
        for j in range(dim):
            m = halves[(c >> j) & 1]
            acc = m if acc is None else np.kron(acc, m)
        assert np.allclose(acc.T, oo.m2m[c], rtol=0, atol=1e-14)
    L.fb_ops_free(h)


OP_CASES = [
    # p, dim, radius, depth, kernel, compression, eps
    (6, 3, 1.001, 4, 0, 2, 1e-6),
    (5, 3, 0.501, 3, 0, 1, 1e-5),
    (6, 3, 0.501, 3, 2, 2, 1e-6),
    (6, 3, 0.501, 3, 3, 2, 1e-6),
    (7, 2, 0.501, 4, 1, 2, 1e-7),
    (8, 2, 0.501, 3, 1, 1, 1e-8),
    (6, 1, 0.501, 4, 7, 2, 1e-6),
    (5, 3, 1.501, 3, 9, 2, 1e-5),
    (4, 3, 0.501, 2, 0, 0, 1e-4),
]


@pytest.mark.parametrize("p,dim,radius,depth,kt,comp,eps", OP_CASES)
def test_m2l_operators_match_oracle(p, dim, radius, depth, kt, comp, eps):
    """Same ranks, and the truncated product U*Vt agrees to round-off (the factors themselves are
    only unique up to sign / rotation inside equal singular values)."""
    L = _lib.lib()
    h = host_ops(p, dim, radius, depth, kt, comp, eps)
    kern = ok.Kernel(kt)
    oo = oc.PrecomputeOperators(p, dim, radius, depth, kern, comp, eps)
    P = p ** dim
    for lvl in range(2, depth + 1):
        cl = radius / 2 ** (lvl - 1)
        for r in range(len(oo.ref_vecs)):
            U, Vt = get_op(h, lvl, r, P, comp != 0)
            K = kern.matrix((oo.ref_vecs[r][None, :] + oo.nodes_nd * 0.5) * cl, oo.nodes_nd * (0.5 * cl))
            nk = np.linalg.norm(K)
            if comp == 0:
                assert np.linalg.norm(U - oo.u[lvl][r]) <= 1e-14 * nk
                continue
            assert L.fb_ops_rank(h, lvl, r) == oo.rank(lvl, r)
            A = U @ Vt
            B = oo.u[lvl][r] @ oo.vt[lvl][r]
            rk = oo.rank(lvl, r)
            sv = np.linalg.svd(K, compute_uv=False)
            # A truncation that cuts through a cluster of equal singular values (symmetric transfer
            # vectors such as [3,3,3] under pure-SVD compression) is not unique: any basis of the
            # cluster gives a valid best rank-k approximation.  Then both must be optimal instead.
            degenerate_cut = comp == 1 and 0 < rk < P and (sv[rk - 1] - sv[rk]) <= 1e-6 * sv[rk - 1]
            if degenerate_cut:
                best = np.sqrt(np.sum(sv[rk:] ** 2))
                assert abs(np.linalg.norm(A - K) - best) <= 1e-9 * nk
                assert abs(np.linalg.norm(B - K) - best) <= 1e-9 * nk
            else:
                assert np.linalg.norm(A - B) <= 1e-11 * nk, (lvl, r)
            assert np.linalg.norm(A - K) <= 20 * eps * nk
    L.fb_ops_free(h)


def test_singular_value_cutoff_rule():
    """aca.rs:210-224: first k whose tail sum of squares drops below eps^2 * total; all kept if never."""
    assert oc.calculate_singular_values_cutoff([1.0, 1e-3, 1e-6], 1e-2) == 1
    assert oc.calculate_singular_values_cutoff([1.0, 1e-3, 1e-6], 1e-4) == 2
    assert oc.calculate_singular_values_cutoff([1.0, 0.5, 0.25], 1e-8) == 3


def test_operator_cache_returns_identical_operators():
    """Operators::build_cached (host_ops.cpp): a second request with the same arguments is served from the process-wide
    cache — bit-identical factors — and any changed argument (radius, epsilon, kernel parameter) misses it."""
    L = _lib.lib()
    P = 6 ** 3
    a = host_ops(6, 3, 0.7310, 3, 3, 2, 1e-6, base_range=0.8)
    b = host_ops(6, 3, 0.7310, 3, 3, 2, 1e-6, base_range=0.8)    # hit
    c = host_ops(6, 3, 0.7310, 3, 3, 2, 1e-6, base_range=0.81)   # another kernel scale
    d = host_ops(6, 3, 0.7311, 3, 3, 2, 1e-6, base_range=0.8)    # another radius
    differs = {"c": False, "d": False}
    for lvl in (2, 3):
        for r in range(4):
            ua, va = get_op(a, lvl, r, P, True)
            ub, vb = get_op(b, lvl, r, P, True)
            assert np.array_equal(ua, ub) and np.array_equal(va, vb)
            for name, h in (("c", c), ("d", d)):
                uo, _ = get_op(h, lvl, r, P, True)
                if uo.shape != ua.shape or not np.array_equal(uo, ua):
                    differs[name] = True
    assert differs["c"] and differs["d"]
    for h in (a, b, c, d):
        L.fb_ops_free(h)

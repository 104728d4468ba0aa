"""CPU tests pinning the oracle restatement of the BBFMM path (no GPU).

The reference holds no golden vectors for the FMM, so the oracle is pinned against
(i) exact dense summation, (ii) the one reproducible FMM test of the reference
(bbfmm.rs:1464-1500, PointOutsideTree{1}), (iii) structural invariants of the interaction lists."""
import numpy as np
import pytest

from oracle import bbfmm as obb
from oracle import chebyshev as oc
from oracle import kernels as ok
from oracle import linear_tree, morton
from tests import helpers as H


@pytest.mark.parametrize("n,dim,kind,kernel,order,comp,adaptive,sparse,bound", [
    (2500, 3, "clustered", 0, 6, 0, True, True, 5e-6),
    (2500, 3, "clustered", 0, 6, 2, True, True, 2e-5),
    (2500, 3, "clustered", 2, 6, 1, True, False, 2e-5),
    (3000, 2, "uniform", 1, 7, 2, True, True, 5e-6),
    (1500, 1, "uniform", 7, 6, 2, True, True, 1e-6),
    (2500, 3, "clustered", 3, 6, 2, False, True, 2e-5),
])
def test_oracle_fmm_matches_dense(n, dim, kind, kernel, order, comp, adaptive, sparse, bound):
    pts = H.make_points(n, dim, kind, seed=2)
    w = np.random.default_rng(3).random((n, 2))
    ot = H.oracle_tree(pts, order, kernel, adaptive, sparse, 20, comp, 10.0 ** -order)
    ot.set_weights(w)
    got = ot.evaluate(w, pts)
    dense = ok.dense_matvec(ok.Kernel(kernel), pts, pts, w)
    assert H.rel_l2(got, dense) < bound


def test_oracle_gradients_match_dense():
    pts = H.make_points(1500, 3, "clustered", seed=4)
    w = np.random.default_rng(5).random((1500, 1))
    for kernel in (0, 2, 3, 7):
        ot = H.oracle_tree(pts, 7, kernel, True, True, 20, 0, 1e-7)
        ot.set_weights(w)
        v, g = ot.evaluate(w, pts, with_gradients=True)
        dv, dg = ok.dense_matvec(ok.Kernel(kernel), pts, pts, w, with_gradients=True)
        assert H.rel_l2(v, dv) < 1e-5
        assert H.rel_l2(g, dg) < 2e-3


def test_point_outside_tree_known_answer():
    """bbfmm.rs:1464-1500"""
    ot = obb.FmmTree(np.array([[0.5]]), 3, ok.Kernel(ok.LINEAR), True, False, [0.0, 1.0])
    w = np.array([[1.0]])
    ot.set_weights(w)
    with pytest.raises(linear_tree.PointOutsideTree) as e:
        ot.evaluate(w, np.array([[0.5], [10.0]]))
    assert e.value.point_index == 1
    assert str(e.value) == "FMM evaluation failed: target point at row 1 lies outside the tree extents"


def test_tree_center_radius_and_keys():
    c, r = morton.calculate_tree_center_and_radius([0.2, 0.3, 0.9, 1.7])   # morton.rs:349-373
    assert c == [0.5, 1.0] and abs(r - 1.001) < 1e-15
    # encode/decode round trip and navigation (morton.rs:58-305)
    k = morton.encode((5, 3, 6), 3, 3)
    assert morton.decode_key(k, 3) == ((5, 3, 6), 3)
    assert morton.get_parent(k, 3) == morton.encode((2, 1, 3), 2, 3)
    assert k in morton.get_children(morton.get_parent(k, 3), 3)
    assert morton.get_child_index(k, 3) == (5 & 1) | ((3 & 1) << 1) | ((6 & 1) << 2)
    assert len(morton.get_neighbours(morton.encode((0, 0, 0), 2, 3), 3)) == 7
    assert len(morton.get_neighbours(morton.encode((1, 1, 1), 2, 3), 3)) == 26


def test_interaction_list_invariants():
    pts = H.make_points(3000, 3, "clustered", seed=9)
    ot = H.oracle_tree(pts, 3, 7, True, True, 15, 0, 1e-3)
    L = ot.lists
    adj = lambda a, b: morton.are_adjacent(a, b, ot.center, ot.radius, 3)
    for leaf, u in L.u_lists.items():
        assert leaf in u and all(c in L.leaves and adj(leaf, c) for c in u)
    for cell, v in L.v_lists.items():
        lvl = cell & morton.LEVEL_MASK
        assert all((c & morton.LEVEL_MASK) == lvl and not adj(cell, c) for c in v)
    for leaf, wl in L.w_lists.items():
        for c in wl:
            assert not adj(leaf, c) and adj(leaf, morton.get_parent(c, 3))
            assert leaf in L.x_lists[c]
    # every (target leaf, source leaf) pair is covered exactly once by U, or through an ancestor pair in V/W/X
    leaves = sorted(L.leaf_source_indices)
    anc = {l: morton.get_ancestors(l, 3) for l in leaves}
    rng = np.random.default_rng(0)
    for _ in range(300):
        a, b = leaves[rng.integers(len(leaves))], leaves[rng.integers(len(leaves))]
        n_u = int(b in L.u_lists.get(a, ()))
        n_v = sum(1 for ca in anc[a] for cb in anc[b] if cb in L.v_lists.get(ca, ()))
        n_w = sum(1 for cb in anc[b] if cb in L.w_lists.get(a, ()))
        n_x = sum(1 for ca in anc[a] if b in L.x_lists.get(ca, ()))
        assert n_u + n_v + n_w + n_x == 1, (a, b, n_u, n_v, n_w, n_x)


def test_m2m_reproduces_polynomials():
    """Chebyshev anterpolation is exact for polynomials of degree < p: M2M then evaluation of a parent expansion
    equals direct child evaluation (chebyshev.rs:146-241)."""
    p = 5
    nodes = oc.generate_chebyshev_nodes(p)
    tn, _ = oc.evaluate_chebyshev_polynomials(p, nodes)
    m2m = oc.get_m2m_transfer_matrices(p, nodes, tn, 1)
    f = lambda x: 1.0 + x - 0.5 * x ** 2 + 0.25 * x ** 4
    for c in range(2):
        child_nodes_in_parent = (nodes + (1 if c else -1)) * 0.5
        # L2L = M2M^T interpolates parent node values to child nodes
        assert np.allclose(m2m[c].T @ f(nodes), f(child_nodes_in_parent), atol=1e-13)


@pytest.mark.parametrize("n,dim,kernel,order,nrhs", [(4000, 3, 0, 5, 2), (3000, 2, 1, 6, 1), (1500, 1, 7, 6, 3)])
def test_c_transfers_match_the_numpy_restatement(n, dim, kernel, order, nrhs):
    """oracle/csrc/oracle_passes.c orc_p2m / orc_transfer_level / orc_l2p (the upward pass, L2L and L2P of the CPU
    baseline) against the per-cell numpy restatement and the pure-Python oracle matvec."""
    from oracle import fast
    from tests import helpers as H
    pts = H.make_points(n, dim, "clustered", seed=3)
    w = np.random.default_rng(1).random((n, nrhs)) - 0.5
    ot = H.oracle_tree(pts, order, kernel, True, True, 30, 2, 1e-6)
    ff = fast.FastFmm(ot)
    M, M2 = ff.upward(w), ff.upward_numpy(w)
    assert np.abs(M - M2).max() <= 1e-13 * np.abs(M2).max()
    ot.set_weights(w)
    assert H.rel_l2(ff.matvec(w), ot.evaluate(w, pts)) <= 1e-13

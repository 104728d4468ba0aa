"""Oracle restatement of ferreus_bbfmm/src/bbfmm.rs FmmTree (test infrastructure only).

FmmTree.new            bbfmm.rs:272-376
set_weights/upward     bbfmm.rs:383-401, 666-772
_eval/downward         bbfmm.rs:444-507, 778-1086
leaf pass              bbfmm.rs:1089-1440
set_local_coefficients bbfmm.rs:518-524, evaluate_leaves bbfmm.rs:537-616
"""
import numpy as np

from . import chebyshev, linear_tree, morton
from .kernels import Kernel


def get_pointarray_extents(points):
    """utils.rs:22-54: [mins..., maxs...]"""
    points = np.asarray(points, dtype=np.float64)
    return list(points.min(axis=0)) + list(points.max(axis=0))


class FmmParams:
    """bbfmm.rs:77-104"""

    def __init__(self, max_points_per_cell=256, compression_type=chebyshev.COMPRESSION_ACA,
                 epsilon=None, eval_chunk_size=1024, interpolation_order=None):
        self.max_points_per_cell = max_points_per_cell
        self.compression_type = compression_type
        self.epsilon = epsilon if epsilon is not None else 10.0 ** (-int(interpolation_order))
        self.eval_chunk_size = eval_chunk_size


class FmmTree:
    def __init__(self, source_points, interpolation_order, kernel: Kernel, adaptive_tree, sparse,
                 extents=None, params=None):
        self.source_points = np.array(source_points, dtype=np.float64)
        if self.source_points.ndim == 1:
            self.source_points = self.source_points[:, None]
        self.p = int(interpolation_order)
        self.kernel = kernel
        self.adaptive_tree = bool(adaptive_tree)
        self.sparse_tree = bool(sparse)
        tree_extents = list(extents) if extents is not None else get_pointarray_extents(self.source_points)
        self.params = params if params is not None else FmmParams(interpolation_order=self.p)
        self.dim = len(tree_extents) // 2
        if self.dim not in (1, 2, 3):
            raise ValueError(f"Unsupported number of dimensions: {self.dim}")
        self.center, self.radius = morton.calculate_tree_center_and_radius(tree_extents)
        self.nrhs = 1
        self.lists = linear_tree.build_tree(self.source_points, self.center, self.radius,
                                            self.params.max_points_per_cell, not self.sparse_tree,
                                            self.dim, self.adaptive_tree)
        self.depth = self.lists.depth
        self.ops = chebyshev.PrecomputeOperators(self.p, self.dim, self.radius, self.depth, self.kernel,
                                                 self.params.compression_type, self.params.epsilon)
        self.P = self.ops.num_nodes_nd
        self.multipoles = {}
        self.locals = {}

    # ------------------------------------------------------------------ helpers
    def _center_length(self, key):
        return morton.get_center_length(key, self.center, self.radius, self.dim)

    def _m2l_index(self, vec):
        """bbfmm.rs:989-998"""
        return sum((7 ** (self.dim - 1 - i)) * (int(v) + 3) for i, v in enumerate(vec))

    # ------------------------------------------------------------------ upward
    def set_weights(self, weights):
        weights = np.asarray(weights, dtype=np.float64)
        if weights.ndim == 1:
            weights = weights[:, None]
        self.nrhs = weights.shape[1]
        t = self.lists
        self.multipoles = {k: np.zeros((self.P, self.nrhs)) for k in t.tree}
        cells_with_sources = set()
        for leaf in t.leaf_source_indices:
            cells_with_sources |= morton.get_ancestors(leaf, self.dim)
        # P2M  bbfmm.rs:691-739
        for key in t.leaves:
            if key in cells_with_sources and key in t.leaf_source_indices:
                idx = t.leaf_source_indices[key]
                c, length = self._center_length(key)
                S, _ = chebyshev.get_approximation_coefficients(self.p, self.source_points[idx], c, length,
                                                                self.ops.polynomial_nodes, self.dim)
                self.multipoles[key] += S.T @ weights[idx]
        # M2M  bbfmm.rs:742-772
        for level in range(self.depth - 1, 0, -1):
            for parent in t.level_cells_map.get(level, []):
                if parent in cells_with_sources:
                    for child in t.children.get(parent, []):
                        ci = morton.get_child_index(child, self.dim)
                        self.multipoles[parent] += self.ops.m2m[ci] @ self.multipoles[child]

    # ------------------------------------------------------------------ downward
    def _downward_pass(self, weights, cells_with_targets):
        t = self.lists
        self.locals = {k: np.zeros((self.P, self.nrhs)) for k in t.tree}
        for level in range(1, self.depth + 1):
            for key in t.level_cells_map.get(level, []):
                if key not in cells_with_targets:
                    continue
                v_list = t.v_lists.get(key)
                if v_list:
                    self._multipole_to_local(key, level, v_list)
                if self.adaptive_tree:
                    x_list = t.x_lists.get(key)
                    if x_list:
                        c, length = self._center_length(key)
                        nodes = chebyshev.scale_cheb_nodes_to_cell(self.ops.nodes_nd, c, length)
                        for xc in x_list:                                   # P2L bbfmm.rs:1001-1048
                            if xc in t.leaf_source_indices:
                                idx = t.leaf_source_indices[xc]
                                a = self.kernel.matrix(nodes, self.source_points[idx])
                                self.locals[key] += a @ weights[idx]
        for level in range(1, self.depth + 1):                               # L2L bbfmm.rs:1051-1086
            for key in t.level_cells_map.get(level, []):
                if key not in cells_with_targets:
                    continue
                for child in t.children.get(key, []):
                    if child in cells_with_targets:
                        ci = morton.get_child_index(child, self.dim)
                        self.locals[child] += self.ops.m2m[ci].T @ self.locals[key]

    def _multipole_to_local(self, key, level, v_list):
        """bbfmm.rs:864-986"""
        c, length = self._center_length(key)
        ops = self.ops
        for v_cell in v_list:
            vc, _ = self._center_length(v_cell)
            vec = [int(round((a - b) / length)) for a, b in zip(c, vc)]
            tidx = self._m2l_index(vec)
            ref = ops.reference_vector_lookups[tidx]
            pidx = ops.permutation_lookups[tidx]
            perm = ops.permutation_indices[pidx]
            inv = ops.inverse_permutations[pidx]
            x = self.multipoles[v_cell][perm, :]
            if ops.compression == chebyshev.COMPRESSION_NONE:
                y = ops.u[level][ref] @ x
            else:
                y = ops.u[level][ref] @ (ops.vt[level][ref] @ x)
            self.locals[key] += y[inv, :]

    # ------------------------------------------------------------------ evaluation
    def _assign_targets(self, target_points):
        keys = linear_tree.points_to_keys(target_points, self.lists.leaves, self.depth, self.center,
                                          self.radius, self.dim)
        self.lists.leaf_target_indices = linear_tree.get_points_to_leaves_map(keys)
        return keys

    def evaluate(self, weights, target_points, with_gradients=False):
        """bbfmm.rs:444-507"""
        weights = np.asarray(weights, dtype=np.float64)
        if weights.ndim == 1:
            weights = weights[:, None]
        target_points = np.asarray(target_points, dtype=np.float64)
        if target_points.ndim == 1:
            target_points = target_points[:, None]
        self._assign_targets(target_points)
        cells_with_targets = set()
        for leaf in self.lists.leaf_target_indices:
            cells_with_targets |= morton.get_ancestors(leaf, self.dim)
        self._downward_pass(weights, cells_with_targets)
        return self._leaf_pass(weights, target_points, with_gradients)

    def set_local_coefficients(self, weights):
        """bbfmm.rs:518-524"""
        weights = np.asarray(weights, dtype=np.float64)
        if weights.ndim == 1:
            weights = weights[:, None]
        self._downward_pass(weights, set(self.lists.tree))

    def evaluate_leaves(self, weights, target_points, with_gradients=False):
        """bbfmm.rs:570-616"""
        weights = np.asarray(weights, dtype=np.float64)
        if weights.ndim == 1:
            weights = weights[:, None]
        target_points = np.asarray(target_points, dtype=np.float64)
        if target_points.ndim == 1:
            target_points = target_points[:, None]
        self._assign_targets(target_points)
        return self._leaf_pass(weights, target_points, with_gradients)

    def _leaf_pass(self, weights, target_points, with_gradients):
        """bbfmm.rs:1113-1440"""
        t = self.lists
        m = target_points.shape[0]
        d = self.dim
        out = np.zeros((m, self.nrhs))
        grads = np.zeros((m, self.nrhs * d)) if with_gradients else None

        def direct(tidx, src_pts, src_w):
            tp = target_points[tidx]
            diffs = [tp[:, k][:, None] - src_pts[:, k][None, :] for k in range(d)]
            r2 = np.zeros((tp.shape[0], src_pts.shape[0]))
            for df in diffs:
                r2 = r2 + df * df
            if with_gradients:
                val, fac = self.kernel.value_and_grad_factor(r2)
                out[tidx] += val @ src_w
                for k in range(d):
                    g = (fac * diffs[k]) @ src_w
                    for j in range(self.nrhs):
                        grads[tidx, j * d + k] += g[:, j]
            else:
                out[tidx] += self.kernel.eval_r2(r2) @ src_w

        for leaf, tidx in t.leaf_target_indices.items():
            u_list = t.u_lists.get(leaf)
            if u_list:                                                        # P2P :1162-1251
                for uc in u_list:
                    if uc in t.leaf_source_indices:
                        sidx = t.leaf_source_indices[uc]
                        direct(tidx, self.source_points[sidx], weights[sidx])
            if self.adaptive_tree:                                            # M2P :1254-1355
                w_list = t.w_lists.get(leaf)
                if w_list:
                    for wc in w_list:
                        c, length = self._center_length(wc)
                        nodes = chebyshev.scale_cheb_nodes_to_cell(self.ops.nodes_nd, c, length)
                        direct(tidx, nodes, self.multipoles[wc])
            c, length = self._center_length(leaf)                             # L2P :1358-1440
            S, dS = chebyshev.get_approximation_coefficients(self.p, target_points[tidx], c, length,
                                                             self.ops.polynomial_nodes, d, with_gradients)
            loc = self.locals[leaf]
            out[tidx] += S @ loc
            if with_gradients:
                for k in range(d):
                    g = dS[:, k * self.P:(k + 1) * self.P] @ loc
                    for j in range(self.nrhs):
                        grads[tidx, j * d + k] += g[:, j]
        return (out, grads) if with_gradients else out

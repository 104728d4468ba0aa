"""Oracle restatement of ferreus_bbfmm/src/linear_tree.rs (test infrastructure only).

build_tree           linear_tree.rs:20-175
adaptive lists       linear_tree.rs:177-395
regular lists        linear_tree.rs:397-485
points_to_keys       linear_tree.rs:487-520
points_to_leaves     linear_tree.rs:522-534
"""
import math
from collections import deque

import numpy as np

from . import morton


class PointOutsideTree(Exception):
    """bbfmm.rs:20-45  FmmError::PointOutsideTree{point_index}."""

    def __init__(self, point_index):
        super().__init__(
            f"FMM evaluation failed: target point at row {point_index} lies outside the tree extents")
        self.point_index = point_index


class TreeLists:
    """bbfmm.rs:111-147"""

    def __init__(self):
        self.tree = set()
        self.leaves = set()
        self.children = {}
        self.u_lists = {}
        self.v_lists = {}
        self.x_lists = None
        self.w_lists = None
        self.level_cells_map = {}
        self.key_to_index_map = {}
        self.leaf_source_indices = {}
        self.leaf_target_indices = {}
        self.depth = 0


def build_tree(points, center, radius, max_points_per_cell, store_empty_leaves, dim, adaptive_tree):
    """Level-synchronous BFS of linear_tree.rs:20-175.

    Binning of a cell's own points into child keys (lines 55-66) is done for all active
    cells of a level at once: side_L = 2r/2^L is an exact power-of-two scaling, hence child
    anchors nest inside parent anchors and grouping the active points by child key gives the
    same per-child index lists (ascending) as the per-cell loop.
    """
    points = np.asarray(points, dtype=np.float64)
    n = points.shape[0]
    displacement = [c - radius for c in center]
    optimal_depth = int(math.ceil(math.log2(float(n)) / float(dim))) if n > 0 else 0

    t = TreeLists()
    t.tree = {0}
    t.level_cells_map = {0: [0]}
    cells_point_indices = {0: np.arange(n, dtype=np.int64)}
    t.key_to_index_map = {0: 0}
    active = [0]
    current_level = 0

    while active:
        next_level = []
        child_level = current_level + 1
        side = morton.get_side_length(radius, child_level)
        any_child_exceeds = False

        # bin all points of active cells at child level
        have = [c for c in active if c in cells_point_indices and len(cells_point_indices[c])]
        if have:
            idx = np.concatenate([cells_point_indices[c] for c in have])
            anchors = morton.points_to_anchors(points[idx], displacement, side)
            keys = morton.encode_anchors(anchors, child_level, dim)
            order = np.argsort(keys, kind="stable")
            sk = keys[order]
            si = idx[order]
            bounds = np.flatnonzero(np.r_[True, sk[1:] != sk[:-1], True])
            for a, b in zip(bounds[:-1], bounds[1:]):
                ck = int(sk[a])
                cells_point_indices[ck] = np.sort(si[a:b])
        children_with_points = {}
        for c in have:
            children_with_points[c] = []
        for ck in [k for k in cells_point_indices if (k & morton.LEVEL_MASK) == child_level]:
            children_with_points[morton.get_parent(ck, dim)].append(ck)

        for cell in active:
            cell_children = sorted(children_with_points.get(cell, []))
            active_children = morton.get_children(cell, dim) if store_empty_leaves else cell_children
            for child in active_children:
                t.tree.add(child)
                t.key_to_index_map.setdefault(child, len(t.tree) - 1)
                t.children.setdefault(child, [])
                t.level_cells_map.setdefault(child_level, []).append(child)
                if child in cells_point_indices:
                    cnt = len(cells_point_indices[child])
                    if adaptive_tree:
                        if cnt > max_points_per_cell and child_level < morton.MAXIMUM_LEVEL:
                            next_level.append(child)
                        else:
                            t.leaves.add(child)
                            t.leaf_source_indices[child] = cells_point_indices[child]
                    elif cnt > max_points_per_cell:
                        any_child_exceeds = True
                elif adaptive_tree and store_empty_leaves:
                    t.leaves.add(child)
            t.children[cell] = list(active_children)
            if not adaptive_tree:
                next_level.extend(active_children)

        should_subdivide = adaptive_tree or (any_child_exceeds
                                             and child_level < morton.MAXIMUM_LEVEL
                                             and child_level < optimal_depth)
        if should_subdivide and next_level:
            active = next_level
            current_level += 1
        else:
            if not adaptive_tree:
                for leaf in next_level:
                    if leaf in cells_point_indices:
                        t.leaf_source_indices.setdefault(leaf, cells_point_indices[leaf])
                t.leaves.update(next_level)
            active = []

    t._cells_point_indices = cells_point_indices
    if adaptive_tree:
        u, v, x, w = get_interaction_lists_adaptive(t.tree, t.leaves, center, radius, dim)
        t.u_lists, t.v_lists, t.x_lists, t.w_lists = u, v, x, w
    else:
        u, v = get_interaction_lists_regular(t.tree, t.leaves, cells_point_indices, t.children,
                                             center, radius, dim)
        t.u_lists, t.v_lists = u, v
    t.depth = current_level + 1
    return t


def get_interaction_lists_adaptive(tree, leaves, center, radius, dim):
    """linear_tree.rs:270-394"""
    u_lists, v_lists, w_lists, x_lists = {}, {}, {}, {}
    geom = {}

    def center_length(k):  # memoised morton.get_center_length
        g = geom.get(k)
        if g is None:
            g = geom[k] = morton.get_center_length(k, center, radius, dim)
        return g

    def adj(a, b):  # morton.rs:308-325
        ca, la = center_length(a)
        cb, lb = center_length(b)
        length = 0.5 * (la + lb)
        for va, vb in zip(ca, cb):
            if not (abs(vb - va) <= 1e-6 + length):
                return False
        return True
    for key in tree:
        u, v, w = set(), set(), set()
        parent = morton.get_parent(key, dim)
        if parent is not None:
            for col in morton.get_neighbours(parent, dim):
                for pcc in morton.get_children(col, dim):
                    if pcc in tree and not adj(key, pcc):
                        v.add(pcc)
            if key in leaves:
                colleagues = morton.get_neighbours(key, dim)
                queue = deque(colleagues)
                visited = set()
                while queue:
                    cur = queue.popleft()
                    if cur in visited:
                        continue
                    visited.add(cur)
                    if adj(key, cur):
                        if cur in leaves:
                            u.add(cur)
                        else:
                            p = morton.get_parent(cur, dim)
                            if p is not None:
                                queue.append(p)
                queue = deque(ch for col in colleagues for ch in morton.get_children(col, dim)
                              if ch in tree)
                while queue:
                    cur = queue.popleft()
                    if adj(key, cur):
                        if cur in leaves:
                            u.add(cur)
                        else:
                            queue.extend(ch for ch in morton.get_children(cur, dim) if ch in tree)
                    else:
                        w.add(cur)
                u.add(key)
        if u:
            u_lists[key] = u
        if v:
            v_lists[key] = v
        if w:
            w_lists[key] = w
    for cell, wl in w_lists.items():
        for wc in wl:
            x_lists.setdefault(wc, set()).add(cell)
    return u_lists, v_lists, x_lists, w_lists


def get_interaction_lists_regular(tree, leaves, cells_points_indices, children, center, radius, dim):
    """linear_tree.rs:397-485"""
    u_lists, v_lists = {}, {}
    for cell in tree:
        u, v = set(), set()
        parent = morton.get_parent(cell, dim)
        if parent is not None:
            if cell in leaves:
                for sib in children.get(parent, []):
                    if sib in cells_points_indices:
                        u.add(sib)
            for pc in morton.get_neighbours(parent, dim):
                if pc not in tree:
                    continue
                for col in children.get(pc, []):
                    if col in cells_points_indices:
                        if morton.are_adjacent(cell, col, center, radius, dim):
                            if cell in leaves:
                                u.add(col)
                        else:
                            v.add(col)
        if cell in leaves:
            u_lists[cell] = u
        v_lists[cell] = v
    return u_lists, v_lists


def points_to_keys(points, leaves, depth, center, radius, dim):
    """linear_tree.rs:487-520: key at level `depth`, walk parents until a leaf; first failing
    index (in row order) is reported."""
    points = np.asarray(points, dtype=np.float64)
    side = morton.get_side_length(radius, depth)
    displacement = [c - radius for c in center]
    anchors = morton.points_to_anchors(points, displacement, side)
    keys = morton.encode_anchors(anchors, depth, dim)
    out = np.empty(points.shape[0], dtype=np.uint64)
    cache = {}
    for i, k in enumerate(keys.tolist()):
        r = cache.get(k, -1)
        if r == -1:
            cur = k
            while cur not in leaves:
                cur = morton.get_parent(cur, dim)
                if cur is None:
                    break
            cache[k] = cur
            r = cur
        if r is None:
            raise PointOutsideTree(i)
        out[i] = r
    return out


def get_points_to_leaves_map(point_keys):
    """linear_tree.rs:522-534"""
    order = np.argsort(point_keys, kind="stable")
    sk = point_keys[order]
    out = {}
    if len(sk) == 0:
        return out
    bounds = np.flatnonzero(np.r_[True, sk[1:] != sk[:-1], True])
    for a, b in zip(bounds[:-1], bounds[1:]):
        out[int(sk[a])] = np.sort(order[a:b])
    return out

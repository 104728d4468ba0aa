"""Oracle restatement of ferreus_bbfmm/src/morton.rs (test infrastructure only).

Keys are ``(bit_interleave(x, y, z; x = LSB) << 15) | level`` with 16 bits per
coordinate (morton.rs:58-119, morton_constants.rs:12-22).  The reference's byte
LUTs are replaced by explicit bit loops; results are identical.
"""
import functools
import math

import numpy as np

MAXIMUM_LEVEL = 16          # morton_constants.rs:12
LEVEL_DISPLACEMENT = 15     # morton_constants.rs:15
LEVEL_MASK = 0x7FFF         # morton_constants.rs:18

# morton_constants.rs:33-77 (order matters only for iteration order, kept for fidelity)
DIRECTIONS = {
    1: [(-1,), (1,)],
    2: [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)],
    3: [(-1, -1, -1), (-1, -1, 0), (-1, -1, 1), (-1, 0, -1), (-1, 1, -1), (-1, 0, 0), (-1, 0, 1),
        (-1, 1, 0), (-1, 1, 1), (0, -1, -1), (1, -1, -1), (0, -1, 0), (0, -1, 1), (1, -1, 0),
        (1, -1, 1), (0, 0, -1), (0, 1, -1), (1, 0, -1), (1, 1, -1), (0, 0, 1), (0, 1, 0),
        (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)],
}


def get_side_length(radius, level):
    """morton.rs:29-32"""
    return 2.0 * radius / float(1 << level)


def calculate_tree_center_and_radius(extents):
    """morton.rs:349-373; extents = [mins..., maxs...]."""
    d = len(extents) // 2
    lo = [math.floor(v) for v in extents[:d]]
    hi = [math.ceil(v) for v in extents[d:]]
    center = [(a + b) / 2.0 for a, b in zip(lo, hi)]
    radius = -math.inf
    for a, b in zip(lo, hi):
        radius = max(radius, (b - a) / 2.0 + 1e-3)
    return center, radius


def points_to_anchors(points, displacement, side_length):
    """Vectorised morton.rs:35-51: floor((x - disp)/side) as u64 (Rust `as` saturates:
    negative/NaN -> 0, >= 2^64 -> u64::MAX)."""
    points = np.asarray(points, dtype=np.float64)
    disp = np.asarray(displacement, dtype=np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        v = np.floor((points - disp[None, :]) / side_length)
    v = np.where(np.isnan(v), 0.0, v)
    big = v >= 18446744073709551616.0
    v = np.clip(v, 0.0, 1.8e19)
    a = v.astype(np.uint64)
    a[big] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return a


def encode_anchors(anchors, level, dim):
    """Vectorised morton.rs:58-119 (coordinates are masked to 16 bits by the byte LUT lookups)."""
    anchors = np.asarray(anchors, dtype=np.uint64)
    code = np.zeros(anchors.shape[0], dtype=np.uint64)
    for bit in range(16):
        for j in range(dim):
            b = (anchors[:, j] >> np.uint64(bit)) & np.uint64(1)
            code |= b << np.uint64(bit * dim + j)
    return (code << np.uint64(LEVEL_DISPLACEMENT)) | np.uint64(level)


def encode(anchor, level, dim):
    code = 0
    for bit in range(16):
        for j in range(dim):
            code |= ((int(anchor[j]) >> bit) & 1) << (bit * dim + j)
    return (code << LEVEL_DISPLACEMENT) | int(level)


def get_level(key):
    return key & LEVEL_MASK


@functools.lru_cache(maxsize=1 << 20)
def decode_key(key, dim):
    """morton.rs:127-167 -> (anchor tuple, level)."""
    level = key & LEVEL_MASK
    code = key >> LEVEL_DISPLACEMENT
    anchor = [0] * dim
    for bit in range(21 if dim == 3 else (28 if dim == 2 else 16)):
        for j in range(dim):
            anchor[j] |= ((code >> (bit * dim + j)) & 1) << bit
    return tuple(anchor), level


def get_parent(key, dim):
    """morton.rs:170-190"""
    level = key & LEVEL_MASK
    if level == 0:
        return None
    return (((key >> LEVEL_DISPLACEMENT) >> dim) << LEVEL_DISPLACEMENT) | (level - 1)


def get_ancestors(key, dim):
    """morton.rs:193-210 (includes the key itself)."""
    out = {key}
    cur = key
    while True:
        p = get_parent(cur, dim)
        if p is None:
            break
        out.add(p)
        cur = p
    return out


def get_neighbours(key, dim):
    """morton.rs:214-263: same-level cells inside [0, 2^level)^d."""
    anchor, level = decode_key(key, dim)
    nmax = 1 << level
    out = []
    for dvec in DIRECTIONS[dim]:
        a = [anchor[j] + dvec[j] for j in range(dim)]
        if all(0 <= v < nmax for v in a):
            out.append(encode(a, level, dim))
    return out


def get_siblings(key, dim):
    """morton.rs:266-285"""
    level = key & LEVEL_MASK
    root = ((key >> LEVEL_DISPLACEMENT) >> dim) << dim
    return [((root | s) << LEVEL_DISPLACEMENT) | level for s in range(1 << dim)]


def get_children(key, dim):
    """morton.rs:288-297"""
    level = key & LEVEL_MASK
    child = (((key >> LEVEL_DISPLACEMENT) << dim) << LEVEL_DISPLACEMENT) | (level + 1)
    return get_siblings(child, dim)


def get_child_index(key, dim):
    """morton.rs:300-305 (bit j = offset along axis j)."""
    return (key >> LEVEL_DISPLACEMENT) & ((1 << dim) - 1)


def get_center_length(key, tree_center, tree_radius, dim):
    """morton.rs:328-346"""
    anchor, level = decode_key(key, dim)
    side = get_side_length(tree_radius, level)
    center = [(float(anchor[j]) + 0.5) * side + (tree_center[j] - tree_radius) for j in range(dim)]
    return center, side


def are_adjacent(a, b, tree_center, tree_radius, dim):
    """morton.rs:308-325 — float test with absolute tolerance 1e-6."""
    ca, la = get_center_length(a, tree_center, tree_radius, dim)
    cb, lb = get_center_length(b, tree_center, tree_radius, dim)
    length = 0.5 * (la + lb)
    return all(abs(vb - va) <= 1e-6 + length for va, vb in zip(ca, cb))

// Host-side tree + interaction lists. See host_tree.h for the reference mapping.
#include "host_tree.h"

#include <algorithm>
#include <cmath>
#include <deque>
#include <memory>
#include <unordered_set>

namespace fb {

namespace {

inline void deinterleave(uint64_t pfx, int dim, uint32_t a[3]) {
  a[0] = a[1] = a[2] = 0;
  for (int bit = 0; bit < 16; ++bit)
    for (int j = 0; j < dim; ++j) a[j] |= (uint32_t)((pfx >> (bit * dim + j)) & 1ull) << bit;
}

}  // namespace
uint64_t interleave(const uint32_t a[3], int dim) {
  uint64_t code = 0;
  for (int bit = 0; bit < 16; ++bit)
    for (int j = 0; j < dim; ++j) code |= (uint64_t)((a[j] >> bit) & 1u) << (bit * dim + j);
  return code;
}
namespace {
struct LP {  // (level, prefix) pair: a cell position that may or may not exist in the tree
  int level;
  uint64_t prefix;
};

}  // namespace

void HostTree::anchor(int c, uint32_t a[3]) const { deinterleave(prefix[c], dim, a); }

static inline void center_of(const HostTree &t, int lvl, uint64_t pfx, double out[3], double &side) {
  uint32_t a[3];
  deinterleave(pfx, t.dim, a);
  side = 2.0 * t.radius / (double)(1ull << lvl);  // morton.rs:29-32
  for (int j = 0; j < t.dim; ++j) out[j] = ((double)a[j] + 0.5) * side + t.disp[j];  // morton.rs:339-343
}

static inline bool adjacent_lp(const HostTree &t, int la, uint64_t pa, int lb, uint64_t pb) {
  double ca[3], cb[3], sa, sb;
  center_of(t, la, pa, ca, sa);
  center_of(t, lb, pb, cb, sb);
  const double length = 0.5 * (sa + sb);
  for (int j = 0; j < t.dim; ++j)
    if (!(std::fabs(cb[j] - ca[j]) <= 1e-6 + length)) return false;  // morton.rs:315-324
  return true;
}

void HostTree::cell_center(int c, double out[3], double &side) const { center_of(*this, level[c], prefix[c], out, side); }

bool HostTree::adjacent(int a, int b) const { return adjacent_lp(*this, level[a], prefix[a], level[b], prefix[b]); }

int HostTree::find(int lvl, uint64_t pfx) const {
  if (lvl < 0 || lvl > depth) return -1;
  const auto b = prefix.begin() + level_ptr[lvl], e = prefix.begin() + level_ptr[lvl + 1];
  auto it = std::lower_bound(b, e, pfx);
  if (it == e || *it != pfx) return -1;
  return (int)(it - prefix.begin());
}

void HostTree::build(const uint64_t *codes, size_t n, int dim_, const double center_[3], double radius_,
                     size_t max_pts, bool store_empty, bool adaptive_) {
  dim = dim_;
  adaptive = adaptive_;
  sparse = !store_empty;
  radius = radius_;
  for (int j = 0; j < 3; ++j) {
    center[j] = j < dim ? center_[j] : 0.0;
    disp[j] = center[j] - radius;  // linear_tree.rs:30
  }
  const int nchild = 1 << dim;
  const uint64_t optimal_depth = (uint64_t)std::ceil(std::log2((double)n) / (double)dim);  // linear_tree.rs:32

  prefix.assign(1, 0);
  level.assign(1, 0);
  parent.assign(1, -1);
  is_leaf.assign(1, 0);
  pt_begin.assign(1, 0);
  pt_end.assign(1, (int32_t)n);
  level_ptr.assign(1, 0);
  level_ptr.push_back(1);

  std::vector<int32_t> active{0};
  int current_level = 0;
  while (!active.empty()) {
    const int child_level = current_level + 1;
    const int shift = dim * (16 - child_level);
    std::vector<int32_t> next;
    bool any_child_exceeds = false;
    for (int32_t cell : active) {
      int32_t cursor = pt_begin[cell];
      const int32_t cend = pt_end[cell];
      for (int s = 0; s < nchild; ++s) {
        const uint64_t cp = (prefix[cell] << dim) | (uint64_t)s;
        const uint64_t *e = std::partition_point(codes + cursor, codes + cend,
                                                 [&](uint64_t c) { return (c >> shift) <= cp; });
        const int32_t b = cursor, en = (int32_t)(e - codes);
        cursor = en;
        const bool has = en > b;
        if (!has && !store_empty) continue;
        const int32_t id = (int32_t)prefix.size();
        prefix.push_back(cp);
        level.push_back(child_level);
        parent.push_back(cell);
        pt_begin.push_back(b);
        pt_end.push_back(en);
        is_leaf.push_back(0);
        if (adaptive) {
          if (has) {
            if ((size_t)(en - b) > max_pts && child_level < 16)  // linear_tree.rs:89-99
              next.push_back(id);
            else
              is_leaf[id] = 1;
          } else {
            is_leaf[id] = 1;  // linear_tree.rs:103-105
          }
        } else {
          if (has && (size_t)(en - b) > max_pts) any_child_exceeds = true;
          next.push_back(id);
        }
      }
    }
    level_ptr.push_back((int32_t)prefix.size());
    const bool should_subdivide =
        adaptive || (any_child_exceeds && child_level < 16 && (uint64_t)child_level < optimal_depth);
    if (should_subdivide && !next.empty()) {
      active.swap(next);
      current_level += 1;
    } else {
      if (!adaptive)
        for (int32_t id : next) is_leaf[id] = 1;  // linear_tree.rs:123-130
      active.clear();
    }
  }
  depth = current_level + 1;
  // level_ptr has depth+2 entries: levels 0..depth
  while ((int)level_ptr.size() < depth + 2) level_ptr.push_back((int32_t)prefix.size());

  const size_t nc = prefix.size();
  // children CSR (cells of a level were appended parent by parent, suffix ascending)
  child_ptr.assign(nc + 1, 0);
  for (size_t c = 1; c < nc; ++c) child_ptr[parent[c] + 1]++;
  for (size_t c = 0; c < nc; ++c) child_ptr[c + 1] += child_ptr[c];
  child_idx.assign(nc > 0 ? nc - 1 : 0, 0);
  {
    std::vector<int32_t> fill(child_ptr.begin(), child_ptr.end() - 1);
    for (size_t c = 1; c < nc; ++c) child_idx[fill[parent[c]]++] = (int32_t)c;
  }
  // leaves in Morton order
  leaves.clear();
  for (size_t c = 0; c < nc; ++c)
    if (is_leaf[c]) leaves.push_back((int32_t)c);
  std::sort(leaves.begin(), leaves.end(), [&](int32_t a, int32_t b) {
    return (prefix[a] << (dim * (16 - level[a]))) < (prefix[b] << (dim * (16 - level[b])));
  });

  if (adaptive)
    build_lists_adaptive();
  else
    build_lists_regular();
}

static void neighbours_of(const HostTree &t, int lvl, uint64_t pfx, std::vector<uint64_t> &out) {
  // morton.rs:214-263 — same-level cells inside [0, 2^level)^d, excluding the cell itself
  out.clear();
  uint32_t a[3];
  deinterleave(pfx, t.dim, a);
  const int64_t nmax = (int64_t)1 << lvl;
  const int dz0 = t.dim > 2 ? -1 : 0, dz1 = t.dim > 2 ? 1 : 0;
  const int dy0 = t.dim > 1 ? -1 : 0, dy1 = t.dim > 1 ? 1 : 0;
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = dy0; dy <= dy1; ++dy)
      for (int dz = dz0; dz <= dz1; ++dz) {
        if (dx == 0 && dy == 0 && dz == 0) continue;
        const int64_t x = (int64_t)a[0] + dx, y = (int64_t)a[1] + dy, z = (int64_t)a[2] + dz;
        if (x < 0 || x >= nmax) continue;
        if (t.dim > 1 && (y < 0 || y >= nmax)) continue;
        if (t.dim > 2 && (z < 0 || z >= nmax)) continue;
        uint32_t b[3] = {(uint32_t)x, (uint32_t)(t.dim > 1 ? y : 0), (uint32_t)(t.dim > 2 ? z : 0)};
        out.push_back(interleave(b, t.dim));
      }
}

static void to_csr(std::vector<std::vector<int32_t>> &lists, std::vector<int64_t> &ptr, std::vector<int32_t> &idx) {
  const size_t nc = lists.size();
  ptr.assign(nc + 1, 0);
  for (size_t c = 0; c < nc; ++c) {
    auto &l = lists[c];
    std::sort(l.begin(), l.end());
    l.erase(std::unique(l.begin(), l.end()), l.end());
    ptr[c + 1] = ptr[c] + (int64_t)l.size();
  }
  idx.resize((size_t)ptr[nc]);
  for (size_t c = 0; c < nc; ++c) std::copy(lists[c].begin(), lists[c].end(), idx.begin() + ptr[c]);
}

void HostTree::build_lists_adaptive() {
  const size_t nc = ncells();
  const int nchild = 1 << dim;
  std::vector<std::vector<int32_t>> U(nc), V(nc), W(nc), X(nc);
#pragma omp parallel
  {
    std::vector<uint64_t> nb, nb2;
    std::deque<LP> queue;
    std::unordered_set<uint64_t> visited;
#pragma omp for schedule(dynamic, 64)
    for (long ci = 1; ci < (long)nc; ++ci) {
      const int c = (int)ci;
      const int lvl = level[c];
      const uint64_t pfx = prefix[c];
      // V list: children of the parent's colleagues, in tree, not adjacent (linear_tree.rs:277-293)
      neighbours_of(*this, lvl - 1, pfx >> dim, nb);
      for (uint64_t np : nb)
        for (int s = 0; s < nchild; ++s) {
          const int id = find(lvl, (np << dim) | (uint64_t)s);
          if (id >= 0 && !adjacent(c, id)) V[c].push_back(id);
        }
      if (!is_leaf[c]) continue;
      // U list, upward sweep through colleagues' ancestors (linear_tree.rs:295-328)
      neighbours_of(*this, lvl, pfx, nb);
      queue.clear();
      visited.clear();
      for (uint64_t np : nb) queue.push_back(LP{lvl, np});
      while (!queue.empty()) {
        const LP cur = queue.front();
        queue.pop_front();
        if (!visited.insert((cur.prefix << 15) | (uint64_t)cur.level).second) continue;
        if (adjacent_lp(*this, lvl, pfx, cur.level, cur.prefix)) {
          const int id = find(cur.level, cur.prefix);
          if (id >= 0 && is_leaf[id])
            U[c].push_back(id);
          else if (cur.level > 0)
            queue.push_back(LP{cur.level - 1, cur.prefix >> dim});
        }
      }
      // downward sweep through colleagues' descendants (linear_tree.rs:330-362)
      queue.clear();
      for (uint64_t np : nb)
        for (int s = 0; s < nchild; ++s) {
          const uint64_t cp = (np << dim) | (uint64_t)s;
          if (find(lvl + 1, cp) >= 0) queue.push_back(LP{lvl + 1, cp});
        }
      while (!queue.empty()) {
        const LP cur = queue.front();
        queue.pop_front();
        const int id = find(cur.level, cur.prefix);
        if (adjacent_lp(*this, lvl, pfx, cur.level, cur.prefix)) {
          if (is_leaf[id]) {
            U[c].push_back(id);
          } else {
            for (int k = child_ptr[id]; k < child_ptr[id + 1]; ++k) {
              const int ch = child_idx[k];
              queue.push_back(LP{level[ch], prefix[ch]});
            }
          }
        } else {
          W[c].push_back(id);
        }
      }
      U[c].push_back(c);
    }
  }
  for (size_t c = 0; c < nc; ++c)
    for (int32_t w : W[c]) X[w].push_back((int32_t)c);  // linear_tree.rs:388-392
  to_csr(U, u_ptr, u_idx);
  to_csr(V, v_ptr, v_idx);
  to_csr(W, w_ptr, w_idx);
  to_csr(X, x_ptr, x_idx);
}

void HostTree::build_lists_regular() {
  const size_t nc = ncells();
  std::vector<std::vector<int32_t>> U(nc), V(nc), E(nc);
  std::vector<uint64_t> nb;
  for (size_t ci = 1; ci < nc; ++ci) {
    const int c = (int)ci;
    const int p = parent[c];
    const bool leaf = is_leaf[c];
    if (leaf)  // linear_tree.rs:453-461
      for (int k = child_ptr[p]; k < child_ptr[p + 1]; ++k) {
        const int sib = child_idx[k];
        if (pt_end[sib] > pt_begin[sib]) U[c].push_back(sib);
      }
    neighbours_of(*this, level[p], prefix[p], nb);
    for (uint64_t np : nb) {
      const int pc = find(level[p], np);
      if (pc < 0) continue;
      for (int k = child_ptr[pc]; k < child_ptr[pc + 1]; ++k) {
        const int col = child_idx[k];
        if (pt_end[col] <= pt_begin[col]) continue;
        if (adjacent(c, col)) {
          if (leaf) U[c].push_back(col);
        } else {
          V[c].push_back(col);
        }
      }
    }
  }
  to_csr(U, u_ptr, u_idx);
  to_csr(V, v_ptr, v_idx);
  to_csr(E, w_ptr, w_idx);
  to_csr(E, x_ptr, x_idx);
}

}  // namespace fb

// ---- host-only tree + interaction lists (no GPU) ----------------------------------------------------------------------
// The product path sorts the level-16 codes on the device (fmm.cu) and hands them to HostTree::build; this entry point
// computes and sorts the same codes on the host, so the CPU test-suite can compare keys, leaf membership and the U / V /
// W / X lists with the oracle bit for bit (morton.rs:29-373, linear_tree.rs:20-485) without a device.
#include "../../include/ferreus_b200.h"

struct fb_host_tree {
  fb::HostTree ht;
  std::vector<uint32_t> perm;  // sorted position -> source row (stable: equal codes keep their row order, like the radix sort)
  size_t n = 0;
};

extern "C" {

int fb_host_tree_new(const double *points, size_t n, int dim, ptrdiff_t row_stride, ptrdiff_t col_stride,
                     const double *extents_or_null, uint64_t max_points_per_cell, int adaptive_tree, int sparse,
                     fb_host_tree **out) {
  if (!out) return FB_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (!points || n == 0 || dim < 1 || dim > 3 || max_points_per_cell == 0) return FB_ERR_INVALID_ARGUMENT;
  try {
    double ext[6];
    if (extents_or_null) {
      for (int d = 0; d < 2 * dim; ++d) ext[d] = extents_or_null[d];
    } else {
      for (int d = 0; d < dim; ++d) ext[d] = ext[dim + d] = points[(ptrdiff_t)d * col_stride];
      for (size_t i = 0; i < n; ++i)
        for (int d = 0; d < dim; ++d) {
          const double v = points[(ptrdiff_t)i * row_stride + (ptrdiff_t)d * col_stride];
          if (v < ext[d]) ext[d] = v;
          if (v > ext[dim + d]) ext[dim + d] = v;
        }
    }
    double center[3] = {0, 0, 0}, radius = -INFINITY;  // calculate_tree_center_and_radius, morton.rs:349-373
    for (int d = 0; d < dim; ++d) {
      const double lo = std::floor(ext[d]), hi = std::ceil(ext[dim + d]);
      center[d] = (lo + hi) / 2.0;
      radius = std::max(radius, (hi - lo) / 2.0 + 1e-3);
    }
    if (!(std::isfinite(radius) && radius > 0)) return FB_ERR_INVALID_ARGUMENT;
    const double side16 = 2.0 * radius / 65536.0;  // morton.rs:29-32 at level 16
    std::vector<uint64_t> codes(n);
    for (size_t i = 0; i < n; ++i) {
      uint32_t a[3] = {0, 0, 0};
      for (int d = 0; d < dim; ++d) {
        const double x = points[(ptrdiff_t)i * row_stride + (ptrdiff_t)d * col_stride];
        const double q = std::floor((x - (center[d] - radius)) / side16);  // point_to_anchor, morton.rs:35-51
        const uint64_t v = !(q >= 0.0) ? 0ull : (q >= 18446744073709551616.0 ? ~0ull : (uint64_t)q);
        if (v >= 65536ull) return FB_ERR_INVALID_ARGUMENT;  // outside the given extents (fmm.cu rejects it the same way)
        a[d] = (uint32_t)v;
      }
      codes[i] = fb::interleave(a, dim);
    }
    std::unique_ptr<fb_host_tree> t(new fb_host_tree());
    t->n = n;
    t->perm.resize(n);
    for (size_t i = 0; i < n; ++i) t->perm[i] = (uint32_t)i;
    std::stable_sort(t->perm.begin(), t->perm.end(), [&](uint32_t x, uint32_t y) { return codes[x] < codes[y]; });
    std::vector<uint64_t> sorted(n);
    for (size_t i = 0; i < n; ++i) sorted[i] = codes[t->perm[i]];
    t->ht.build(sorted.data(), n, dim, center, radius, (size_t)max_points_per_cell, !sparse, adaptive_tree != 0);
    *out = t.release();
    return FB_OK;
  } catch (...) {
    return FB_ERR_INVALID_ARGUMENT;
  }
}

void fb_host_tree_free(fb_host_tree *t) { delete t; }

int fb_host_tree_counts(const fb_host_tree *t, uint64_t *n_cells, uint64_t *n_leaves, int32_t *depth, uint64_t *n_list4) {
  if (!t) return FB_ERR_INVALID_ARGUMENT;
  if (n_cells) *n_cells = t->ht.ncells();
  if (n_leaves) *n_leaves = t->ht.leaves.size();
  if (depth) *depth = t->ht.depth;
  if (n_list4) {
    n_list4[0] = t->ht.u_idx.size();
    n_list4[1] = t->ht.v_idx.size();
    n_list4[2] = t->ht.w_idx.size();
    n_list4[3] = t->ht.x_idx.size();
  }
  return FB_OK;
}

int fb_host_tree_dump_cells(const fb_host_tree *t, uint64_t *keys, uint8_t *leaf_flags, uint64_t *leaf_ptr,
                            uint64_t *leaf_idx) {
  if (!t) return FB_ERR_INVALID_ARGUMENT;
  const size_t nc = t->ht.ncells();
  uint64_t off = 0;
  for (size_t c = 0; c < nc; ++c) {
    if (keys) keys[c] = t->ht.ref_key((int)c);
    if (leaf_flags) leaf_flags[c] = t->ht.is_leaf[c];
    if (leaf_ptr) leaf_ptr[c] = off;
    if (t->ht.is_leaf[c]) {
      const int b = t->ht.pt_begin[c], e = t->ht.pt_end[c];
      if (leaf_idx) {
        for (int i = b; i < e; ++i) leaf_idx[off + (i - b)] = t->perm[i];
        std::sort(leaf_idx + off, leaf_idx + off + (e - b));
      }
      off += (uint64_t)(e - b);
    }
  }
  if (leaf_ptr) leaf_ptr[nc] = off;
  return FB_OK;
}

int fb_host_tree_dump_list(const fb_host_tree *t, int which, uint64_t *ptr, uint64_t *idx) {
  if (!t || which < 0 || which > 3) return FB_ERR_INVALID_ARGUMENT;
  const std::vector<int64_t> *p[4] = {&t->ht.u_ptr, &t->ht.v_ptr, &t->ht.w_ptr, &t->ht.x_ptr};
  const std::vector<int32_t> *x[4] = {&t->ht.u_idx, &t->ht.v_idx, &t->ht.w_idx, &t->ht.x_idx};
  if (ptr)
    for (size_t i = 0; i < p[which]->size(); ++i) ptr[i] = (uint64_t)(*p[which])[i];
  if (idx)
    for (size_t i = 0; i < x[which]->size(); ++i) idx[i] = (uint64_t)(*x[which])[i];
  return FB_OK;
}

}  // extern "C"

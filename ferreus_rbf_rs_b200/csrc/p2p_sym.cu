// Symmetric P2P for the matvec case targets == sources, one right-hand side (the solver's and the bench's hot call).
// In a partitioned tree (comm.cu) a rank runs it for the chunks of its own Morton range [own_lo, own_hi) only: a chunk
// still takes ALL sources behind it, whoever owns them, and the source-side sums of foreign rows travel in the result
// all-reduce — every unordered pair of the whole cloud is evaluated exactly once across the ranks, nothing twice.
// Reference: particle_to_particle bbfmm.rs:1162-1251 — there every (target, source) pair of the U lists is evaluated;
// here each unordered pair is evaluated ONCE and serves both rows:
//     out[t] += k(t, s) w[s]      and      out[s] += k(t, s) w[t]        (all registry kernels are symmetric).
// U is a symmetric relation (adjacent leaves at any level + the leaf itself, linear_tree.rs:177-395 / :397-485), and the
// sources are stored in Morton order, so "each unordered pair once" = a warp owning the 32 consecutive sorted positions
// [tb, tb + cnt) of a leaf takes the part of the leaf's merged U ranges that lies AFTER its own chunk (symmetric
// tiles), plus its own cnt x cnt diagonal block evaluated in full.
//
// Cost model (tools/fp64_ubench.cu, tools/dmma_bench.cu): the FP64 pipe is the only resource (DMMA shares it, MUFU.RSQ64H
// serialises with it), 10 DFMA + 1 MUFU per evaluation = 7.2 clk per warp step on an SM.  The symmetric step adds one
// DMUL (k w[t]) and, amortised, one DADD of the source-side reduction: 8.3 clk for TWO pairs — 1.75x fewer pipe cycles
// per pair.  Source side: lane l parks k(t_l, s_j) w[t_l] in a warp-private shared-memory matrix part[j][l] (row stride
// 33: conflict-free both ways); after the tile, lane j sums row j in a fixed order and issues one RED to out[s_j].
// Target-side sums stay in registers and are added with one RED per target at the end (other warps' source-side REDs
// hit the same rows concurrently).  Values are the reference's own per-pair values; only the summation order differs.
#include "fmm.h"

#include <algorithm>
#include <cstdlib>

namespace fb {

constexpr int kSymWPC = 4;       // warps per CTA
constexpr int kSymStride = 33;   // row stride of the partial-sum matrix (doubles)

// NR right-hand sides per pass (1, 2 or 4): the kernel value of a pair is computed once and serves 2 NR accumulations.
// The source-side partial sums need NR matrices, so tiles shrink to 16 / 8 sources for NR = 2 / 4 (shared memory per
// warp: 10.5 / 11.4 / 12.5 KB; with 16-source tiles NR = 4 fitted two CTAs per SM and gained 9 % on the 2-D thin-plate
// config where the model says 40 %).
template <int NR>
struct SymWarpSmem {
  static constexpr int TS = NR == 1 ? 32 : (NR == 2 ? 16 : 8);  // sources per tile
  static constexpr int NV = 2 + NR / 2;         // double2 planes per source: {x, y}, {z, w0}, {w1, w2}, {w3, -}
  double2 st[2][NV][32];                        // double-buffered source tile (32 slots: the diagonal block uses them all)
  double part[NR * TS * kSymStride];            // part[r][j][l] = sum over the lane's targets of k(t, s_j) w_r[t]
};

template <int NR>
__device__ __forceinline__ void sym_unpack_w(const double2 (*t)[32], int j, double z_w0_y, double (&ws)[NR]) {
  ws[0] = z_w0_y;
  if (NR >= 2) {
    const double2 p2 = t[2][j];
    ws[1] = p2.x;
    if (NR >= 3) ws[2] = p2.y;
  }
  if (NR >= 4) ws[3] = t[3][j].x;
}

// A warp owns up to 64 consecutive sorted positions of a leaf, lane l the targets tb + l and (TWO) tb + 32 + l: every
// staged source then serves two evaluations per lane, which halves the shared-memory traffic per pair (source broadcasts,
// partial-sum stores and the row sums of the flush; ncu on the one-target version: l1tex 70 % busy, FP64 pipe 48 %) and
// doubles the independent chains in flight.  Chunks of <= 32 targets take the one-target instantiation.
// split > 1 (grids too small to fill the GPU, e.g. one rank's share): `split` warps share a chunk, warp `part` takes
// every split-th source tile (part 0 also the diagonal block); all sums leave through REDs, so nothing else changes.
template <int FAM, bool FAST, bool TWO, int NR>
__device__ __forceinline__ void p2p_sym_body(const DirectArgs &a, SymWarpSmem<NR> &sm, const int li, const int tb,
                                             const int cnt, const int lane, const int split, const int part) {
  constexpr int NT = TWO ? 2 : 1;
  constexpr int TS = SymWarpSmem<NR>::TS;
  constexpr double kScale = kernel_weight_scale<FAM, FAST>();
  const int a_end = tb + cnt;
  double xt[NT], yt[NT], zt[NT], wt[NT][NR], acc[NT][NR];
#pragma unroll
  for (int u = 0; u < NT; ++u) {
    const int i = 32 * u + lane;
    xt[u] = yt[u] = zt[u] = 0.0;
#pragma unroll
    for (int r = 0; r < NR; ++r) wt[u][r] = acc[u][r] = 0.0;
    if (i < cnt) {
      xt[u] = a.sx[tb + i];
      yt[u] = a.sy[tb + i];
      zt[u] = a.sz[tb + i];
#pragma unroll
      for (int r = 0; r < NR; ++r) wt[u][r] = a.w[(size_t)(a.rhs0 + r) * a.n + tb + i] * kScale;
    }
  }
  auto park = [&](int buf, double x, double y, double z, const double (&w)[NR]) {  // this lane's source slot of a tile
    sm.st[buf][0][lane] = make_double2(x, y);
    sm.st[buf][1][lane] = make_double2(z, w[0]);
    if (NR >= 2) sm.st[buf][2][lane] = make_double2(w[1], NR >= 3 ? w[2] : 0.0);
    if (NR >= 4) sm.st[buf][3][lane] = make_double2(w[3], 0.0);
  };
  // ---- diagonal block: the chunk against itself, every ordered pair (self term included, as the reference does)
#pragma unroll
  for (int h = 0; h < NT; ++h) park(h, xt[h], yt[h], zt[h], wt[h]);
  __syncwarp();
#pragma unroll
  for (int h = 0; h < NT; ++h) {
    const int mh = part == 0 ? min(32, cnt - 32 * h) : 0;
#pragma unroll 2
    for (int j = 0; j < mh; ++j) {
      const double2 p0 = sm.st[h][0][j], p1 = sm.st[h][1][j];
      double ws[NR];
      sym_unpack_w<NR>(sm.st[h], j, p1.y, ws);
#pragma unroll
      for (int u = 0; u < NT; ++u) {
        const double dx = xt[u] - p0.x, dy = yt[u] - p0.y, dz = zt[u] - p1.x;
        double r2 = dx * dx;
        r2 += dy * dy;
        r2 += dz * dz;
        const double v = kernel_mag<FAM, FAST, true>(r2, a.kp);
#pragma unroll
        for (int r = 0; r < NR; ++r) kernel_acc<FAM>(acc[u][r], v, ws[r]);
      }
    }
  }
  __syncwarp();

  // ---- symmetric tiles: the part of the merged U ranges behind the chunk
  long long e = a.u_ptr[li];
  const long long e_end = a.u_ptr[li + 1];
  int pos = 0, end = 0;  // current clipped range [pos, end)
  auto next_range = [&]() {
    while (e < e_end) {
      const int b = a.u_begin[e], n = a.u_count[e];
      ++e;
      pos = max(b, a_end);
      end = b + n;
      if (pos < end) return;
    }
    pos = end = 0;
  };
  next_range();
  double rx = 0, ry = 0, rz = 0, rw[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) rw[r] = 0.0;
  int base_next = 0;
  int skip = part;  // tiles to pass over before the next one of this warp
  auto fetch = [&](int &m) {  // this lane's element of the next tile of this warp, then advance
    for (;;) {
      m = min(TS, end - pos);
      if (m <= 0) {
        m = 0;
        return;
      }
      const bool mine = skip == 0;
      skip = mine ? split - 1 : skip - 1;
      if (mine) {
        base_next = pos;
        if (lane < m) {
          rx = a.sx[pos + lane];
          ry = a.sy[pos + lane];
          rz = a.sz[pos + lane];
#pragma unroll
          for (int r = 0; r < NR; ++r) rw[r] = a.w[(size_t)(a.rhs0 + r) * a.n + pos + lane] * kScale;
        }
      }
      pos += m;
      if (pos >= end) next_range();
      if (mine) return;
    }
  };
  int m_cur = 0, m_next = 0, buf = 0, base_cur = 0;
  fetch(m_cur);
  base_cur = base_next;
  if (lane < m_cur) park(0, rx, ry, rz, rw);
  __syncwarp();
  while (m_cur > 0) {
    fetch(m_next);  // global loads of the next tile overlap the arithmetic below
    const double2(*t)[32] = sm.st[buf];
    double *pl = sm.part + lane;
#pragma unroll 4
    for (int j = 0; j < m_cur; ++j) {
      const double2 p0 = t[0][j], p1 = t[1][j];
      double ws[NR], ps[NR];
      sym_unpack_w<NR>(t, j, p1.y, ws);
#pragma unroll
      for (int u = 0; u < NT; ++u) {
        const double dx = xt[u] - p0.x, dy = yt[u] - p0.y, dz = zt[u] - p1.x;
        double r2 = dx * dx;
        r2 += dy * dy;
        r2 += dz * dz;
        const double v = kernel_mag<FAM, FAST, true>(r2, a.kp);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          kernel_acc<FAM>(acc[u][r], v, ws[r]);
          ps[r] = u == 0 ? v * wt[0][r] : fma(v, wt[u][r], ps[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) pl[(r * TS + j) * kSymStride] = ps[r];
    }
    __syncwarp();
    // source side: row (r, j) of the partial sums, fixed order, one RED per (source, right-hand side)
    for (int q = lane; q < m_cur * NR; q += 32) {
      const int r = q / m_cur, j = q - r * m_cur;
      const double *pr = sm.part + (r * TS + j) * kSymStride;
      double sp[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) sp[k] = pr[k];
#pragma unroll
      for (int l = 8; l < 32; l += 8)
#pragma unroll
        for (int k = 0; k < 8; ++k) sp[k] += pr[l + k];
      const double s = ((sp[0] + sp[1]) + (sp[2] + sp[3])) + ((sp[4] + sp[5]) + (sp[6] + sp[7]));
      atomicAdd(a.out + (size_t)a.sym_row[base_cur + j] * a.nrhs + a.rhs0 + r, FAM == KF_LINEAR ? -s : s);
    }
    buf ^= 1;
    if (lane < m_next) park(buf, rx, ry, rz, rw);
    __syncwarp();
    m_cur = m_next;
    base_cur = base_next;
  }
#pragma unroll
  for (int u = 0; u < NT; ++u) {
    const int i = 32 * u + lane;
    if (i < cnt)
#pragma unroll
      for (int r = 0; r < NR; ++r) atomicAdd(a.out + (size_t)a.sym_row[tb + i] * a.nrhs + a.rhs0 + r, acc[u][r]);
  }
}

template <int FAM, bool FAST, int NR>
__global__ void __launch_bounds__(kSymWPC * 32, NR == 1 ? 5 : (NR == 2 ? 4 : 3)) k_p2p_sym(const DirectArgs a, const int split) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * kSymWPC + warp;
  const long long chunk = gw / split;
  const int part = (int)(gw - chunk * split);
  const int tile = (int)(chunk >> 1), sub = (int)(chunk & 1);  // a tile holds <= kTile = 128 targets: two warps of <= 64
  if (tile >= *a.ts.n_tiles_dev) return;
  const int li = a.ts.tile_leaf[tile];
  // target i of the set is the source at the sorted position own_lo + i
  const int tb = a.ts.own_lo + a.ts.leaf_begin[li] + a.ts.tile_off[tile] + sub * 64;
  const int cnt = min(64, a.ts.own_lo + a.ts.leaf_end[li] - tb);
  if (cnt <= 0) return;
  extern __shared__ __align__(16) unsigned char dsm_raw[];
  SymWarpSmem<NR> &sm = reinterpret_cast<SymWarpSmem<NR> *>(dsm_raw)[warp];
  if (cnt > 32) p2p_sym_body<FAM, FAST, true, NR>(a, sm, li, tb, cnt, lane, split, part);
  else p2p_sym_body<FAM, FAST, false, NR>(a, sm, li, tb, cnt, lane, split, part);
}

template <int FAM, int NR>
static void p2p_sym_launch(const DirectArgs &a, cudaStream_t s) {
  static_assert(kTile == 128, "two 64-target warps per tile");
  const size_t smem = sizeof(SymWarpSmem<NR>) * kSymWPC;
  // about three waves of warps (148 SMs x 20 resident warps) keep the uneven chunks from leaving SMs idle at the end
  const long long chunks = std::max<long long>(1, (long long)(a.ts.m / 64));
  const int split = (int)std::min<long long>(8, std::max<long long>(1, (148 * 20 * 3) / chunks));
  const unsigned grid = (unsigned)(((long long)a.ts.max_tiles * 2 * split + kSymWPC - 1) / kSymWPC);
  if (kernel_has_fast<FAM>() && a.kp.fast) {
    constexpr bool F = kernel_has_fast<FAM>();
    FB_CUDA(cudaFuncSetAttribute(k_p2p_sym<FAM, F, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FB_LAUNCH((k_p2p_sym<FAM, F, NR>), grid, kSymWPC * 32, smem, s, a, split);
  } else {
    FB_CUDA(cudaFuncSetAttribute(k_p2p_sym<FAM, false, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FB_LAUNCH((k_p2p_sym<FAM, false, NR>), grid, kSymWPC * 32, smem, s, a, split);
  }
}

// 1 .. 4 right-hand sides: one pass of 4, 2 or 1 columns at a time (3 = 2 + 1)
template <int FAM>
static void p2p_sym_fam(DirectArgs a, cudaStream_t s) {
  const int r_end = a.rhs0 + a.nrhs_pass;
  while (a.rhs0 < r_end) {
    const int left = r_end - a.rhs0;
    if (left >= 4) {
      p2p_sym_launch<FAM, 4>(a, s);
      a.rhs0 += 4;
    } else if (left >= 2) {
      p2p_sym_launch<FAM, 2>(a, s);
      a.rhs0 += 2;
    } else {
      p2p_sym_launch<FAM, 1>(a, s);
      a.rhs0 += 1;
    }
  }
}

// true when the symmetric kernel serves this call: values only, at most four right-hand sides (beyond that the kernel
// value is cheap next to the 2 K accumulations and the one-pass general kernel wins), targets = the tree's own sources
// or a rank's Morton-contiguous share of them
bool p2p_sym_applicable(const DirectArgs &a) {
  static const bool off = [] {
    const char *v = std::getenv("FB_P2P_SYM");
    return v && v[0] == '0';
  }();
  return !off && !a.gout && a.nrhs >= 1 && a.nrhs <= 4 && a.ts.own_hi > a.ts.own_lo && a.sym_row != nullptr && a.ts.max_tiles > 0 &&
         a.kp.fast != 3;
}

void launch_p2p_sym(const DirectArgs &a0, cudaStream_t s) {
  DirectArgs a = a0;
  a.rhs0 = 0;
  a.nrhs_pass = a.nrhs;
#define CALL(F) p2p_sym_fam<F>(a, s)
  FB_FAM_SWITCH(a.kp.fam, CALL)
#undef CALL
}

}  // namespace fb

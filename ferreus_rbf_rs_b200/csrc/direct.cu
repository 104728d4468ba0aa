// Direct-sum kernels of the leaf pass (FP64 FMA-pipe bound): P2P + M2P for general target sets, several right-hand
// sides and gradients.  The matvec case (targets == sources, one right-hand side) takes the symmetric P2P kernel of
// p2p_sym.cu; the downward-pass P2L (and its fused M2P transpose) lives in p2l.cu; p2p_mma.cu is an opt-in experiment.
// Reference: particle_to_particle bbfmm.rs:1162-1251, multipole_to_particle :1254-1355.
//   k_leaf_warp   values: one warp = 32 targets of a leaf, warp-private source tiles, the kernel function evaluated
//                 once per pair for all right-hand sides (the hot path);
//   k_leaf_direct values + gradients (evaluate_with_gradients): one thread per target, CTA-wide source tiles of kTile
//                 points (or the Chebyshev nodes of a W cell, generated on the fly) broadcast from shared memory.
#include "fmm.h"

#include <algorithm>
#include <cstdlib>

namespace fb {

template <int NR>
struct SrcTile {
  double x[kTile], y[kTile], z[kTile];
  double w[NR][kTile];
};

template <int FAM, int NR, bool GRAD>
__device__ __forceinline__ void accumulate_tile(const SrcTile<NR> &t, int m, double xt, double yt, double zt,
                                                const KParams &kp, double (&acc)[NR], double (&g)[GRAD ? 3 : 1]) {
#pragma unroll 4
  for (int j = 0; j < m; ++j) {
    const double dx = xt - t.x[j], dy = yt - t.y[j], dz = zt - t.z[j];
    double r2 = dx * dx;
    r2 += dy * dy;
    r2 += dz * dz;
    if (GRAD) {
      double v, f;
      kernel_value_grad<FAM>(r2, kp, v, f);
      const double w0 = t.w[0][j];
      acc[0] += v * w0;
      const double fw = f * w0;
      g[0] += fw * dx;
      g[1] += fw * dy;
      g[2] += fw * dz;
    } else {
      const double v = kernel_mag<FAM, false, true>(r2, kp);
#pragma unroll
      for (int r = 0; r < NR; ++r) kernel_acc<FAM>(acc[r], v, t.w[r][j]);
    }
  }
}

template <int FAM, int NR, bool GRAD>
__global__ void __launch_bounds__(kTile) k_leaf_direct(const DirectArgs a) {
  const int tile = blockIdx.x;
  if (tile >= *a.ts.n_tiles_dev) return;
  const int li = a.ts.tile_leaf[tile];
  const int tb = a.ts.leaf_begin[li] + a.ts.tile_off[tile];
  const int cnt = min(kTile, a.ts.leaf_end[li] - tb);
  const int tid = threadIdx.x;
  const bool active = tid < cnt;
  double xt = 0, yt = 0, zt = 0;
  if (active) {
    xt = a.ts.x[tb + tid];
    yt = a.ts.y[tb + tid];
    zt = a.ts.z[tb + tid];
  }
  double acc[NR];
  double g[GRAD ? 3 : 1];
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r] = 0.0;
  g[0] = 0.0;
  if (GRAD) g[1] = g[2] = 0.0;

  __shared__ SrcTile<NR> st;

  // ---- P2P over the merged source ranges of the U list
  for (long long e = a.u_ptr[li]; e < a.u_ptr[li + 1]; ++e) {
    const int b = a.u_begin[e], n = a.u_count[e];
    for (int c0 = 0; c0 < n; c0 += kTile) {
      const int m = min(kTile, n - c0);
      __syncthreads();
      if (tid < m) {
        const int s = b + c0 + tid;
        st.x[tid] = a.sx[s];
        st.y[tid] = a.sy[s];
        st.z[tid] = a.sz[s];
#pragma unroll
        for (int r = 0; r < NR; ++r) st.w[r][tid] = a.w[(size_t)(a.rhs0 + r) * a.n + s];
      }
      __syncthreads();
      accumulate_tile<FAM, NR, GRAD>(st, m, xt, yt, zt, a.kp, acc, g);
    }
  }
  // ---- M2P over the Chebyshev nodes of the W cells
  const int p = a.p, P = a.P;
  for (long long e = a.w_ptr[li]; e < a.w_ptr[li + 1]; ++e) {
    const int c = a.w_cell[e];
    const double cx = a.ccx[c], cy = a.ccy[c], cz = a.ccz[c], h = a.chalf[c];
    for (int c0 = 0; c0 < P; c0 += kTile) {
      const int m = min(kTile, P - c0);
      __syncthreads();
      if (tid < m) {
        const int nd = c0 + tid;
        int i0, i1, i2;
        if (a.dim == 3) {
          i2 = nd % p;
          i1 = (nd / p) % p;
          i0 = nd / (p * p);
        } else if (a.dim == 2) {
          i1 = nd % p;
          i0 = nd / p;
          i2 = 0;
        } else {
          i0 = nd;
          i1 = i2 = 0;
        }
        st.x[tid] = cx + h * a.nodes[i0];  // chebyshev.rs:951-968
        st.y[tid] = a.dim > 1 ? cy + h * a.nodes[i1] : 0.0;
        st.z[tid] = a.dim > 2 ? cz + h * a.nodes[i2] : 0.0;
#pragma unroll
        for (int r = 0; r < NR; ++r) st.w[r][tid] = a.mult[((size_t)c * a.nrhs + a.rhs0 + r) * coef_stride(P) + nd];
      }
      __syncthreads();
      accumulate_tile<FAM, NR, GRAD>(st, m, xt, yt, zt, a.kp, acc, g);
    }
  }
  if (active) {
    const size_t row = a.ts.out_row[tb + tid];
#pragma unroll
    for (int r = 0; r < NR; ++r) a.out[row * a.nrhs + a.rhs0 + r] += acc[r];
    if (GRAD) {
      double *gp = a.gout + row * (size_t)(a.nrhs * a.dim) + (size_t)a.rhs0 * a.dim;
      for (int d = 0; d < a.dim; ++d) gp[d] += g[d];
    }
  }
}

constexpr int kWarpTile = 32;
constexpr int kRegOrder = 8;

// ======================================================================================================
// Leaf pass, warp-granular (value only).  ncu on the earlier CTA-granular kernel (profiles/r1_s2_ncu_full.txt):
// FP64 pipe 58 % busy, 41 % warps active — a 128-thread CTA serving a 30-point leaf parks three idle warps on
// the M2P barriers.  Here the work item is one warp = 32 targets of one leaf: no block barrier anywhere, warps
// without targets retire at once, CTAs are two warps so the block scheduler balances uneven leaves.
//   P2P  warp-private double-buffered 32-source tiles: the next tile is prefetched into registers while the
//        current one is consumed, only __syncwarp;
//   M2P  squared axis offsets per (target, W cell) in a warp-private table, multipoles staged in warp-private
//        chunks of whole i0-slabs, r^2 = (dx2[i0] + dy2[i1]) + dz2[i2] (utils.rs:230-237 association); W cells
//        are never adjacent to the leaf, so r^2 > 0 and the zero-distance select is dropped.
// ======================================================================================================
constexpr int kLeafWPC = 2;         // warps per CTA
constexpr int kLeafMwDoubles = 512;  // multipole staging budget per warp (doubles)
constexpr int kM2PQB = 4;            // (i0, i1) node columns in flight per lane

struct LeafWarpCfg {  // per-warp shared-memory layout (doubles) and M2P chunking
  int slabs_per_chunk, plane, d2, total;
};
static inline LeafWarpCfg leaf_warp_cfg(int nr, int p, int dim, bool regz) {
  const int p1 = dim > 1 ? p : 1, p2 = dim > 2 ? p : 1;
  auto plane_of = [&](int slabs) { return ((slabs * p1 + kM2PQB - 1) / kM2PQB) * kM2PQB * p2; };
  LeafWarpCfg c;
  c.slabs_per_chunk = p;
  while (c.slabs_per_chunk > 1 && nr * plane_of(c.slabs_per_chunk) > kLeafMwDoubles) --c.slabs_per_chunk;
  c.plane = plane_of(c.slabs_per_chunk);  // node columns padded to whole groups of kM2PQB (zero multipoles)
  c.d2 = (regz ? 2 : 3) * p * 32;
  const int p2p = 2 * (4 + nr) * kWarpTile;  // double-buffered {x,y},{z,w0},{w1,w2}... pairs
  c.total = std::max(p2p, nr * c.plane + c.d2);
  c.total = (c.total + 1) & ~1;  // keep every warp's slice 16-byte aligned
  return c;
}

template <int FAM, int NR, bool REGZ, bool FAST>
__global__ void __launch_bounds__(kLeafWPC * 32, NR >= 4 ? 7 : 10) k_leaf_warp(const DirectArgs a, const int slabs_per_chunk,
                                                                 const int plane, const int warp_doubles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * kLeafWPC + warp;
  const int tile = (int)(gw >> 2), sub = (int)(gw & 3);
  if (tile >= *a.ts.n_tiles_dev) return;
  const int li = a.ts.tile_leaf[tile];
  const int tb = a.ts.leaf_begin[li] + a.ts.tile_off[tile] + sub * 32;
  const int cnt = min(32, a.ts.leaf_end[li] - tb);
  if (cnt <= 0) return;
  const bool active = lane < cnt;
  double xt = 0, yt = 0, zt = 0;
  if (active) {
    xt = a.ts.x[tb + lane];
    yt = a.ts.y[tb + lane];
    zt = a.ts.z[tb + lane];
  }
  double acc[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r] = 0.0;

  extern __shared__ __align__(16) double dsm[];
  double *wsm = dsm + (size_t)warp * warp_doubles;

  // ---- P2P: warp-private tiles over the merged U ranges; a tile row is {x,y},{z,w0},{w1,w2},... so a source costs
  //      two 128-bit broadcast loads (one more per further pair of right-hand sides)
  if (!a.skip_p2p) {
    constexpr int NC2 = (4 + NR) / 2;  // double2 components per source
    double2(*wt)[NC2][kWarpTile] = reinterpret_cast<double2(*)[NC2][kWarpTile]>(wsm);
    long long e = a.u_ptr[li];
    const long long e_end = a.u_ptr[li + 1];
    int rb = 0, rn = 0, c0 = 0;
    if (e < e_end) {
      rb = a.u_begin[e];
      rn = a.u_count[e];
    }
    double reg[2 * NC2];
#pragma unroll
    for (int k = 0; k < 2 * NC2; ++k) reg[k] = 0.0;
    auto fetch = [&](int &m) {  // this lane's element of the current tile, then advance
      m = 0;
      if (e >= e_end) return;
      m = min(kWarpTile, rn - c0);
      if (lane < m) {
        const int s = rb + c0 + lane;
        reg[0] = a.sx[s];
        reg[1] = a.sy[s];
        reg[2] = a.sz[s];
#pragma unroll
        for (int r = 0; r < NR; ++r) reg[3 + r] = a.w[(size_t)(a.rhs0 + r) * a.n + s] * kernel_weight_scale<FAM, FAST>();
      }
      c0 += kWarpTile;
      if (c0 >= rn) {
        ++e;
        c0 = 0;
        if (e < e_end) {
          rb = a.u_begin[e];
          rn = a.u_count[e];
        }
      }
    };
    auto stash = [&](int buf, int m) {
      if (lane < m) {
#pragma unroll
        for (int k = 0; k < NC2; ++k) wt[buf][k][lane] = make_double2(reg[2 * k], reg[2 * k + 1]);
      }
    };
    int m_cur = 0, m_next = 0, buf = 0;
    fetch(m_cur);
    stash(0, m_cur);
    __syncwarp();
    while (m_cur > 0) {
      fetch(m_next);  // global loads of the next tile overlap the arithmetic below
      const double2(*t)[kWarpTile] = wt[buf];
#pragma unroll 4
      for (int j = 0; j < m_cur; ++j) {
        double sv[2 * NC2];
#pragma unroll
        for (int k = 0; k < NC2; ++k) {
          const double2 v2 = t[k][j];
          sv[2 * k] = v2.x;
          sv[2 * k + 1] = v2.y;
        }
        const double dx = xt - sv[0], dy = yt - sv[1], dz = zt - sv[2];
        double r2 = dx * dx;
        r2 += dy * dy;
        r2 += dz * dz;
        const double v = kernel_mag<FAM, FAST, true>(r2, a.kp);
#pragma unroll
        for (int r = 0; r < NR; ++r) kernel_acc<FAM>(acc[r], v, sv[3 + r]);
      }
      buf ^= 1;
      stash(buf, m_next);
      __syncwarp();
      m_cur = m_next;
    }
  }
  // ---- M2P: tensor grid of the W cell's Chebyshev nodes
  const int p = a.p, P = a.P;
  const int p1 = a.dim > 1 ? p : 1, p2 = a.dim > 2 ? p : 1;
  const int slab = p1 * p2;  // nodes sharing one i0
  double *mw = wsm;          // [NR][plane]: multipoles of the staged i0 slabs, zero padded to whole column groups
  double *d2 = wsm + NR * plane;  // [2 or 3][p][32]
  for (long long e = a.w_ptr[li]; e < a.w_ptr[li + 1]; ++e) {
    const int c = a.w_cell[e];
    const double h = a.chalf[c];
    const double ccx = a.ccx[c], ccy = a.ccy[c], ccz = a.ccz[c];
    double dzr[kRegOrder];
#pragma unroll
    for (int i = 0; i < kRegOrder; ++i) dzr[i] = 1.0;
    __syncwarp();
    for (int i = 0; i < p; ++i) {
      const double nd = a.nodes[i];
      const double ox = xt - (ccx + h * nd);
      d2[(0 * p + i) * 32 + lane] = ox * ox;
      if (a.dim > 1) {
        const double oy = yt - (ccy + h * nd);
        d2[(1 * p + i) * 32 + lane] = oy * oy;
      }
      if (a.dim > 2 && !REGZ) {
        const double oz = zt - (ccz + h * nd);
        d2[(2 * p + i) * 32 + lane] = oz * oz;
      }
    }
    if (REGZ && a.dim > 2) {
#pragma unroll
      for (int i = 0; i < kRegOrder; ++i)
        if (i < p) {
          const double oz = zt - (ccz + h * a.nodes[i]);
          dzr[i] = oz * oz;
        }
    }
    const double *msrc = a.mult + ((size_t)c * a.nrhs + a.rhs0) * coef_stride(P);
    for (int s0 = 0; s0 < p; s0 += slabs_per_chunk) {
      const int ns = min(slabs_per_chunk, p - s0);
      const int cn = ns * slab;  // nodes in this chunk
      const int nq = ns * p1;    // (i0, i1) columns in this chunk
      __syncwarp();
#pragma unroll
      for (int r = 0; r < NR; ++r)
        for (int k = lane; k < plane; k += 32)
          mw[r * plane + k] = k < cn ? msrc[(size_t)r * coef_stride(P) + s0 * slab + k] * kernel_weight_scale<FAM, FAST>() : 0.0;
      __syncwarp();
      // kM2PQB columns at a time: independent kernel evaluations in flight inside every (uniformly predicated) i2
      // step; columns past nq re-read the last column's offsets against zero multipoles
      int csi = s0, ci1 = 0;
      for (int q0 = 0; q0 < nq; q0 += kM2PQB) {
        double axy[kM2PQB];
        const double *wrow[kM2PQB];
#pragma unroll
        for (int u = 0; u < kM2PQB; ++u) {
          const double ax = d2[(0 * p + csi) * 32 + lane];
          axy[u] = a.dim > 1 ? ax + d2[(1 * p + ci1) * 32 + lane] : ax;
          wrow[u] = mw + (q0 + u) * p2;
          if (q0 + u + 1 < nq && ++ci1 == p1) {
            ci1 = 0;
            ++csi;
          }
        }
        if (REGZ) {
#pragma unroll
          for (int i2 = 0; i2 < kRegOrder; ++i2)
            if (i2 < p2) {
              double v[kM2PQB];
#pragma unroll
              for (int u = 0; u < kM2PQB; ++u) v[u] = kernel_mag<FAM, FAST, false>(axy[u] + dzr[i2], a.kp);
#pragma unroll
              for (int u = 0; u < kM2PQB; ++u)
#pragma unroll
                for (int r = 0; r < NR; ++r) kernel_acc<FAM>(acc[r], v[u], wrow[u][r * plane + i2]);
            }
        } else {
          for (int i2 = 0; i2 < p2; ++i2) {
            const double dz2 = a.dim > 2 ? d2[(2 * p + i2) * 32 + lane] : 0.0;
            double v[kM2PQB];
#pragma unroll
            for (int u = 0; u < kM2PQB; ++u) v[u] = kernel_mag<FAM, FAST, false>(a.dim > 2 ? axy[u] + dz2 : axy[u], a.kp);
#pragma unroll
            for (int u = 0; u < kM2PQB; ++u)
#pragma unroll
              for (int r = 0; r < NR; ++r) kernel_acc<FAM>(acc[r], v[u], wrow[u][r * plane + i2]);
          }
        }
      }
    }
  }
  if (active) {
    const size_t row = a.ts.out_row[tb + lane];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double *o = a.out + row * a.nrhs + a.rhs0 + r;
      if (a.atomic_out) atomicAdd(o, acc[r]);
      else *o += acc[r];
    }
  }
}

template <int FAM, int NR, bool REGZ, bool FAST>
static void launch_leaf_warp(const DirectArgs &a, const LeafWarpCfg &cfg, cudaStream_t s) {
  const size_t smem = sizeof(double) * (size_t)cfg.total * kLeafWPC;
  const unsigned grid = (unsigned)(((long long)a.ts.max_tiles * 4 + kLeafWPC - 1) / kLeafWPC);
  if (smem > 48 * 1024)
    FB_CUDA(cudaFuncSetAttribute(k_leaf_warp<FAM, NR, REGZ, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  FB_LAUNCH((k_leaf_warp<FAM, NR, REGZ, FAST>), grid, kLeafWPC * 32, smem, s, a, cfg.slabs_per_chunk, cfg.plane, cfg.total);
}

template <int FAM, int NR>
static void launch_leaf_v3(const DirectArgs &a, cudaStream_t s) {
  const bool regz = a.dim == 3 && a.p <= kRegOrder;
  const LeafWarpCfg cfg = leaf_warp_cfg(NR, a.p, a.dim, regz);
  if (kernel_has_fast<FAM>() && a.kp.fast) {
    if (regz) launch_leaf_warp<FAM, NR, true, kernel_has_fast<FAM>()>(a, cfg, s);
    else launch_leaf_warp<FAM, NR, false, kernel_has_fast<FAM>()>(a, cfg, s);
  } else {
    if (regz) launch_leaf_warp<FAM, NR, true, false>(a, cfg, s);
    else launch_leaf_warp<FAM, NR, false, false>(a, cfg, s);
  }
}

// ---------------------------------------------------------------------------------- dispatch
template <int FAM>
static void leaf_direct_fam(DirectArgs a, cudaStream_t s) {
  const int grid = a.ts.max_tiles;
  if (grid <= 0) return;
  if (a.gout) {
    for (int r = 0; r < a.nrhs; ++r) {
      a.rhs0 = r;
      FB_LAUNCH((k_leaf_direct<FAM, 1, true>), grid, kTile, 0, s, a);
    }
    return;
  }
  if (a.skip_p2p) {  // W lists only
    if (!a.has_w) return;
  } else if (p2p_sym_applicable(a)) {  // targets == sources, 1 RHS: each unordered pair once; the warp kernel keeps the W lists
    launch_p2p_sym(a, s);
    if (!a.has_w) return;
    a.skip_p2p = 1;
  } else if (p2p_mma_applicable(a)) {  // opt-in experiment: squared distances on the FP64 tensor cores
    launch_p2p_mma(a, s);
    if (!a.has_w) return;
    a.skip_p2p = 1;
  }
  int r = 0;
  while (r < a.nrhs) {
    a.rhs0 = r;
    const int left = a.nrhs - r;
    if (left >= 8) {
      launch_leaf_v3<FAM, 8>(a, s);
      r += 8;
    } else if (left >= 4) {
      launch_leaf_v3<FAM, 4>(a, s);
      r += 4;
    } else if (left >= 2) {
      launch_leaf_v3<FAM, 2>(a, s);
      r += 2;
    } else {
      launch_leaf_v3<FAM, 1>(a, s);
      r += 1;
    }
  }
}

void launch_leaf_direct(const DirectArgs &a, cudaStream_t s) {
#define CALL(F) leaf_direct_fam<F>(a, s)
  FB_FAM_SWITCH(a.kp.fam, CALL)
#undef CALL
}

}  // namespace fb

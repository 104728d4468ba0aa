// Direct-sum kernels (FP64 FMA-pipe bound): the leaf pass P2P + M2P and the downward-pass P2L.
// Reference: particle_to_particle bbfmm.rs:1162-1251, multipole_to_particle :1254-1355,
// particle_to_local :1001-1048.  One thread owns one target; source tiles of kTile points (or
// Chebyshev nodes of a W cell, generated on the fly) are staged in shared memory and read back as
// warp-wide broadcasts; the kernel function is evaluated once per pair for all right-hand sides.
#include "fmm.h"

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
      const double v = kernel_value_dev<FAM>(r2, kp);
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[r] += v * t.w[r][j];
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
        for (int r = 0; r < NR; ++r) st.w[r][tid] = a.mult[((size_t)c * a.nrhs + a.rhs0 + r) * P + nd];
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

// P2L: targets are the Chebyshev nodes of a cell, sources the points of its X-list leaves.
template <int FAM, int NR>
__global__ void __launch_bounds__(kTile) k_p2l(const P2LArgs a) {
  const int ci = blockIdx.x;
  const int c = a.cells[ci];
  if (!a.cell_flag[c]) return;
  const int tid = threadIdx.x;
  const int p = a.p, P = a.P;
  const int nd = blockIdx.y * kTile + tid;
  const bool active = nd < P;
  double xt = 0, yt = 0, zt = 0;
  if (active) {
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
    const double h = a.chalf[c];
    xt = a.ccx[c] + h * a.nodes[i0];
    yt = a.dim > 1 ? a.ccy[c] + h * a.nodes[i1] : 0.0;
    zt = a.dim > 2 ? a.ccz[c] + h * a.nodes[i2] : 0.0;
  }
  double acc[NR];
  double g[1] = {0.0};
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r] = 0.0;
  __shared__ SrcTile<NR> st;
  for (long long e = a.x_ptr[ci]; e < a.x_ptr[ci + 1]; ++e) {
    const int b = a.x_begin[e], n = a.x_count[e];
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
      accumulate_tile<FAM, NR, false>(st, m, xt, yt, zt, a.kp, acc, g);
    }
  }
  if (active) {
#pragma unroll
    for (int r = 0; r < NR; ++r) a.loc[((size_t)c * a.nrhs + a.rhs0 + r) * P + nd] += acc[r];
  }
}


// ======================================================================================================
// Value-only fast path of the leaf pass.  Measured on B200 (profiles/): the one-target-per-lane kernels are
// FP64-pipe bound per issued warp; what is lost is lane utilisation and operation count, so:
//  * P2P: every warp streams its own 32-source tiles through a private, double-buffered shared-memory slot
//    (next tile prefetched into registers while the current one is consumed; only __syncwarp, no block barrier),
//    and warps without targets skip the loop entirely;
//  * M2P: the sources are a tensor grid of Chebyshev nodes, so r^2 = (dx2[i0] + dy2[i1]) + dz2[i2] with the
//    squared axis offsets computed once per (target, W cell): 1 + 7 FP64 operations per pair instead of 13, with
//    the reference's association (ferreus_rbf_utils/src/utils.rs:230-237).  For p <= 8 the dz2 table lives in
//    registers (uniformly predicated unroll), otherwise in shared memory.
// ======================================================================================================
constexpr int kWarpTile = 32;
constexpr int kRegOrder = 8;

template <int NR>
struct WarpTile {  // [buffer][component][lane]
  double v[2][3 + NR][kWarpTile];
};

template <int FAM, int NR, bool REGZ>
__global__ void __launch_bounds__(kTile) k_leaf_direct_v2(const DirectArgs a) {
  const int tile = blockIdx.x;
  if (tile >= *a.ts.n_tiles_dev) return;
  const int li = a.ts.tile_leaf[tile];
  const int tb = a.ts.leaf_begin[li] + a.ts.tile_off[tile];
  const int cnt = min(kTile, a.ts.leaf_end[li] - tb);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool active = tid < cnt;
  double xt = 0, yt = 0, zt = 0;
  if (active) {
    xt = a.ts.x[tb + tid];
    yt = a.ts.y[tb + tid];
    zt = a.ts.z[tb + tid];
  }
  double acc[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r] = 0.0;

  extern __shared__ double dsm[];
  WarpTile<NR> *wt = reinterpret_cast<WarpTile<NR> *>(dsm) + warp;
  double *mw = dsm + (sizeof(WarpTile<NR>) / sizeof(double)) * (kTile / 32);  // [NR][P] multipoles of the W cell
  double *d2 = mw + (size_t)NR * a.P;                                          // [2 or 3][p][kTile] squared offsets
  const bool warp_has_targets = warp * 32 < cnt;

  // ---- P2P: warp-private tiles over the merged U ranges
  if (warp_has_targets) {
    long long e = a.u_ptr[li];
    const long long e_end = a.u_ptr[li + 1];
    int rb = 0, rn = 0, c0 = 0;
    if (e < e_end) {
      rb = a.u_begin[e];
      rn = a.u_count[e];
    }
    double reg[3 + NR];
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
        for (int r = 0; r < NR; ++r) reg[3 + r] = a.w[(size_t)(a.rhs0 + r) * a.n + s];
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
    int m_cur = 0, m_next = 0, buf = 0;
    fetch(m_cur);
    if (lane < m_cur) {
#pragma unroll
      for (int k = 0; k < 3 + NR; ++k) wt->v[0][k][lane] = reg[k];
    }
    __syncwarp();
    while (m_cur > 0) {
      fetch(m_next);  // global loads of the next tile overlap the arithmetic below
      const double(*t)[kWarpTile] = wt->v[buf];
#pragma unroll 4
      for (int j = 0; j < m_cur; ++j) {
        const double dx = xt - t[0][j], dy = yt - t[1][j], dz = zt - t[2][j];
        double r2 = dx * dx;
        r2 += dy * dy;
        r2 += dz * dz;
        const double v = kernel_value_dev<FAM>(r2, a.kp);
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r] += v * t[3 + r][j];
      }
      buf ^= 1;
      if (lane < m_next) {
#pragma unroll
        for (int k = 0; k < 3 + NR; ++k) wt->v[buf][k][lane] = reg[k];
      }
      __syncwarp();
      m_cur = m_next;
    }
  }
  // ---- M2P: tensor grid of the W cell's Chebyshev nodes
  const int p = a.p, P = a.P;
  const int p1 = a.dim > 1 ? p : 1, p2 = a.dim > 2 ? p : 1;
  for (long long e = a.w_ptr[li]; e < a.w_ptr[li + 1]; ++e) {
    const int c = a.w_cell[e];
    const double h = a.chalf[c];
    __syncthreads();
    for (int i = tid; i < NR * P; i += kTile) mw[i] = a.mult[((size_t)c * a.nrhs + a.rhs0) * P + i];
    double dzr[kRegOrder];
#pragma unroll
    for (int i = 0; i < kRegOrder; ++i) dzr[i] = 0.0;
    for (int i = 0; i < p; ++i) {
      const double nd = a.nodes[i];
      const double ox = xt - (a.ccx[c] + h * nd);
      d2[(0 * p + i) * kTile + tid] = ox * ox;
      if (a.dim > 1) {
        const double oy = yt - (a.ccy[c] + h * nd);
        d2[(1 * p + i) * kTile + tid] = oy * oy;
      }
      if (a.dim > 2 && !REGZ) {
        const double oz = zt - (a.ccz[c] + h * nd);
        d2[(2 * p + i) * kTile + tid] = oz * oz;
      }
    }
    if (REGZ && a.dim > 2) {
#pragma unroll
      for (int i = 0; i < kRegOrder; ++i)
        if (i < p) {
          const double oz = zt - (a.ccz[c] + h * a.nodes[i]);
          dzr[i] = oz * oz;
        }
    }
    __syncthreads();
    if (!warp_has_targets) continue;
    for (int i0 = 0; i0 < p; ++i0) {
      const double ax = d2[(0 * p + i0) * kTile + tid];
      for (int i1 = 0; i1 < p1; ++i1) {
        const double axy = a.dim > 1 ? ax + d2[(1 * p + i1) * kTile + tid] : ax;
        const double *wrow = mw + (i0 * p1 + i1) * p2;
        if (REGZ) {
#pragma unroll
          for (int i2 = 0; i2 < kRegOrder; ++i2)
            if (i2 < p2) {
              const double r2 = axy + dzr[i2];
              const double v = kernel_value_dev<FAM>(r2, a.kp);
#pragma unroll
              for (int r = 0; r < NR; ++r) acc[r] += v * wrow[(size_t)r * P + i2];
            }
        } else {
#pragma unroll 4
          for (int i2 = 0; i2 < p2; ++i2) {
            const double r2 = a.dim > 2 ? axy + d2[(2 * p + i2) * kTile + tid] : axy;
            const double v = kernel_value_dev<FAM>(r2, a.kp);
#pragma unroll
            for (int r = 0; r < NR; ++r) acc[r] += v * wrow[(size_t)r * P + i2];
          }
        }
      }
    }
  }
  if (active) {
    const size_t row = a.ts.out_row[tb + tid];
#pragma unroll
    for (int r = 0; r < NR; ++r) a.out[row * a.nrhs + a.rhs0 + r] += acc[r];
  }
}

template <int FAM, int NR>
static void launch_leaf_v2(const DirectArgs &a, cudaStream_t s) {
  const bool regz = a.dim == 3 && a.p <= kRegOrder;
  const size_t smem = sizeof(WarpTile<NR>) * (kTile / 32) +
                      sizeof(double) * ((size_t)NR * a.P + (regz ? 2 : 3) * (size_t)a.p * kTile);
  if (regz) {
    if (smem > 48 * 1024)
      FB_CUDA(cudaFuncSetAttribute(k_leaf_direct_v2<FAM, NR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FB_LAUNCH((k_leaf_direct_v2<FAM, NR, true>), a.ts.max_tiles, kTile, smem, s, a);
  } else {
    if (smem > 48 * 1024)
      FB_CUDA(cudaFuncSetAttribute(k_leaf_direct_v2<FAM, NR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FB_LAUNCH((k_leaf_direct_v2<FAM, NR, false>), a.ts.max_tiles, kTile, smem, s, a);
  }
}

template <int FAM, int NR>
static void launch_p2l_v2(const P2LArgs &a, cudaStream_t s) {
  dim3 grid(a.n_cells, (a.P + kTile - 1) / kTile);
  FB_LAUNCH((k_p2l<FAM, NR>), grid, kTile, 0, s, a);
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
  int r = 0;
  while (r < a.nrhs) {
    a.rhs0 = r;
    const int left = a.nrhs - r;
    if (left >= 8) {
      launch_leaf_v2<FAM, 8>(a, s);
      r += 8;
    } else if (left >= 4) {
      launch_leaf_v2<FAM, 4>(a, s);
      r += 4;
    } else if (left >= 2) {
      launch_leaf_v2<FAM, 2>(a, s);
      r += 2;
    } else {
      launch_leaf_v2<FAM, 1>(a, s);
      r += 1;
    }
  }
}

template <int FAM>
static void p2l_fam(P2LArgs a, cudaStream_t s) {
  if (a.n_cells <= 0) return;
  int r = 0;
  while (r < a.nrhs) {
    a.rhs0 = r;
    const int left = a.nrhs - r;
    if (left >= 8) {
      launch_p2l_v2<FAM, 8>(a, s);
      r += 8;
    } else if (left >= 4) {
      launch_p2l_v2<FAM, 4>(a, s);
      r += 4;
    } else if (left >= 2) {
      launch_p2l_v2<FAM, 2>(a, s);
      r += 2;
    } else {
      launch_p2l_v2<FAM, 1>(a, s);
      r += 1;
    }
  }
}

#define FB_FAM_SWITCH(fam, CALL)                         \
  switch (fam) {                                         \
    case KF_LINEAR: CALL(KF_LINEAR); break;              \
    case KF_TPS: CALL(KF_TPS); break;                    \
    case KF_CUBIC: CALL(KF_CUBIC); break;                \
    case KF_SPH: CALL(KF_SPH); break;                    \
    case KF_LAPLACE: CALL(KF_LAPLACE); break;            \
    case KF_R2: CALL(KF_R2); break;                      \
    default: CALL(KF_R4); break;                         \
  }

void launch_leaf_direct(const DirectArgs &a, cudaStream_t s) {
#define CALL(F) leaf_direct_fam<F>(a, s)
  FB_FAM_SWITCH(a.kp.fam, CALL)
#undef CALL
}

void launch_p2l(const P2LArgs &a, cudaStream_t s) {
#define CALL(F) p2l_fam<F>(a, s)
  FB_FAM_SWITCH(a.kp.fam, CALL)
#undef CALL
}

}  // namespace fb

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
      FB_LAUNCH((k_leaf_direct<FAM, 8, false>), grid, kTile, 0, s, a);
      r += 8;
    } else if (left >= 4) {
      FB_LAUNCH((k_leaf_direct<FAM, 4, false>), grid, kTile, 0, s, a);
      r += 4;
    } else if (left >= 2) {
      FB_LAUNCH((k_leaf_direct<FAM, 2, false>), grid, kTile, 0, s, a);
      r += 2;
    } else {
      FB_LAUNCH((k_leaf_direct<FAM, 1, false>), grid, kTile, 0, s, a);
      r += 1;
    }
  }
}

template <int FAM>
static void p2l_fam(P2LArgs a, cudaStream_t s) {
  if (a.n_cells <= 0) return;
  dim3 grid(a.n_cells, (a.P + kTile - 1) / kTile);
  int r = 0;
  while (r < a.nrhs) {
    a.rhs0 = r;
    const int left = a.nrhs - r;
    if (left >= 8) {
      FB_LAUNCH((k_p2l<FAM, 8>), grid, kTile, 0, s, a);
      r += 8;
    } else if (left >= 4) {
      FB_LAUNCH((k_p2l<FAM, 4>), grid, kTile, 0, s, a);
      r += 4;
    } else if (left >= 2) {
      FB_LAUNCH((k_p2l<FAM, 2>), grid, kTile, 0, s, a);
      r += 2;
    } else {
      FB_LAUNCH((k_p2l<FAM, 1>), grid, kTile, 0, s, a);
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

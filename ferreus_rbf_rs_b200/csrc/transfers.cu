// P2M and L2P with the interpolation order and dimension as template parameters (values only).
// Reference: particle_to_multipole bbfmm.rs:691-739, local_to_particle bbfmm.rs:1358-1440,
// get_approximation_coefficients chebyshev.rs:831-927, S_n chebyshev.rs:114-127.
//
// The generic kernels (fmm_kernels.cuh: k_p2m, k_l2p; runtime p and dim) are instruction-bound, not HBM-bound: ncu at
// 1M points, p = 7 (profiles/r1_final_ncu_full.txt) shows 265 M / 220 M warp instructions, l1tex 83 % / 58 % busy, DRAM
// 2 % — ~25 instructions per (point, node) product in P2M (three shared-memory reads, index arithmetic and a global weight
// read per FMA) and ~14 per FMA in L2P (loop control and addressing of runtime-bound loops).  Here
//   P2M  a thread owns one (i0, i1) node column and a residue class of the chunk's points; per point it forms
//        S0[i0] S1[i1] w once and runs the p FMAs of the column against the broadcast S2 row, accumulators in registers
//        across the whole leaf; the point groups are summed through shared memory at the end;
//   L2P  a thread owns one target, its 1-D Chebyshev weights live in registers and the three contractions are fully
//        unrolled: one broadcast shared-memory read and one DFMA per local coefficient.
// Orders without an instantiation fall back to the generic kernels (launch_* returns false).
#include "fmm.h"

#include <cstdlib>

namespace fb {

template <int PO>
__device__ __forceinline__ void cheb_weights(double x, const double *tn, double (&S)[PO]) {
  // S[m] = (2 sum_k T_k(x) T_k(x_m) - 1) / p, k ascending as in the generic kernel; T_k by the three-term recurrence,
  // one k at a time so only the p running sums stay live
  double tkm2 = 1.0, tkm1 = x;
#pragma unroll
  for (int m = 0; m < PO; ++m) S[m] = 0.0;
#pragma unroll
  for (int k = 0; k < PO; ++k) {
    double tk;
    if (k == 0) tk = 1.0;
    else if (k == 1) tk = x;
    else {
      tk = 2.0 * x * tkm1 - tkm2;
      tkm2 = tkm1;
      tkm1 = tk;
    }
#pragma unroll
    for (int m = 0; m < PO; ++m) S[m] += tk * tn[m * PO + k];
    asm volatile("" ::: "memory");  // keep the loads of later k from being hoisted (register pressure)
  }
  constexpr double inv_p = 1.0 / (double)PO;
#pragma unroll
  for (int m = 0; m < PO; ++m) S[m] = (S[m] * 2.0 - 1.0) * inv_p;  // chebyshev.rs:114-127 divides by p: <= 1 ulp apart
}

constexpr int kP2MTChunk = 64;

template <int PO, int DIM>
__global__ void __launch_bounds__(256, PO <= 8 ? 3 : 2) k_p2m_t(const int *leaves, const int *ptb, const int *pte, const double *sx,
                                               const double *sy, const double *sz, const double *w, size_t n,
                                               const double *ccx, const double *ccy, const double *ccz,
                                               const double *chalf, const double *tnodes, int nrhs, double *mult) {
  constexpr int P1 = DIM > 1 ? PO : 1, P2 = DIM > 2 ? PO : 1;
  constexpr int NQ = PO * P1, P = NQ * P2;
  constexpr int QP = NQ <= 32 ? 32 : (NQ <= 64 ? 64 : (NQ <= 128 ? 128 : 256));  // threads per point group
  constexpr int NG = 256 / QP;
  static_assert(NQ <= 256, "order too large for one column per thread");
  __shared__ double tn[PO * PO];
  __shared__ double S[3][kP2MTChunk][PO];
  __shared__ double wS[kP2MTChunk];
  __shared__ double red[NG > 1 ? NG - 1 : 1][P];
  const int c = leaves[blockIdx.x];
  const int b = ptb[c], e = pte[c];
  const int tid = threadIdx.x;
  for (int i = tid; i < PO * PO; i += 256) tn[i] = tnodes[i];
  const double cc0 = ccx[c], cc1 = ccy[c], cc2 = ccz[c];
  const double half = chalf[c];
  const int q = tid % QP, g = tid / QP;
  const int i0 = q / P1, i1 = q % P1;
  for (int r = 0; r < nrhs; ++r) {
    double acc[P2];
#pragma unroll
    for (int i = 0; i < P2; ++i) acc[i] = 0.0;
    for (int c0 = b; c0 < e; c0 += kP2MTChunk) {
      const int m = min(kP2MTChunk, e - c0);
      __syncthreads();
      for (int t = tid; t < m * DIM; t += 256) {
        const int pt = t % m, d = t / m;
        const double coord = d == 0 ? sx[c0 + pt] : (d == 1 ? sy[c0 + pt] : sz[c0 + pt]);
        const double x = (coord - (d == 0 ? cc0 : (d == 1 ? cc1 : cc2))) / half;  // chebyshev.rs:841-845
        double Sv[PO];
        cheb_weights<PO>(x, tn, Sv);
#pragma unroll
        for (int i = 0; i < PO; ++i) S[d][pt][i] = Sv[i];
      }
      if (tid < m) wS[tid] = w[(size_t)r * n + c0 + tid];
      __syncthreads();
      if (q < NQ) {
        for (int pt = g; pt < m; pt += NG) {
          double s01 = S[0][pt][i0] * wS[pt];
          if (DIM > 1) s01 *= S[1][pt][i1];
          if (DIM > 2) {
#pragma unroll
            for (int i2 = 0; i2 < P2; ++i2) acc[i2] += s01 * S[2][pt][i2];
          } else {
            acc[0] += s01;
          }
        }
      }
    }
    // sum the point groups: groups 1.. park their columns, group 0 adds them in a fixed order and writes the leaf
    __syncthreads();
    if (NG > 1 && g > 0 && q < NQ) {
#pragma unroll
      for (int i2 = 0; i2 < P2; ++i2) red[g - 1][q * P2 + i2] = acc[i2];
    }
    __syncthreads();
    if (g == 0 && q < NQ) {
      double *dst = mult + ((size_t)c * nrhs + r) * coef_stride(P) + q * P2;
#pragma unroll
      for (int i2 = 0; i2 < P2; ++i2) {
        double s = acc[i2];
        for (int gg = 1; gg < NG; ++gg) s += red[gg - 1][q * P2 + i2];
        dst[i2] += s;
      }
    }
  }
}

template <int PO, int DIM>
__global__ void __launch_bounds__(kTile, PO <= 8 ? 6 : 4) k_l2p_t(const TargetSet ts, const int *leaf_cell, const double *loc,
                                                 const double *ccx, const double *ccy, const double *ccz,
                                                 const double *chalf, const double *tnodes, int nrhs, double *out) {
  constexpr int P1 = DIM > 1 ? PO : 1, P2 = DIM > 2 ? PO : 1;
  constexpr int P = PO * P1 * P2;
  __shared__ double tn[PO * PO];
  __shared__ double L[P];
  const int tile = blockIdx.x;
  if (tile >= *ts.n_tiles_dev) return;
  const int li = ts.tile_leaf[tile];
  const int tb = ts.leaf_begin[li] + ts.tile_off[tile];
  const int cnt = min(kTile, ts.leaf_end[li] - tb);
  const int tid = threadIdx.x;
  const int c = leaf_cell[li];
  for (int i = tid; i < PO * PO; i += kTile) tn[i] = tnodes[i];
  __syncthreads();
  const bool active = tid < cnt;
  double S0[PO], S1[P1], S2[P2];
  size_t row = 0;
  if (active) {
    const double half = chalf[c];
    cheb_weights<PO>((ts.x[tb + tid] - ccx[c]) / half, tn, S0);
    if (DIM > 1) {
      double t[PO];
      cheb_weights<PO>((ts.y[tb + tid] - ccy[c]) / half, tn, t);
#pragma unroll
      for (int i = 0; i < P1; ++i) S1[i] = t[i];
    } else {
      S1[0] = 1.0;
    }
    if (DIM > 2) {
      double t[PO];
      cheb_weights<PO>((ts.z[tb + tid] - ccz[c]) / half, tn, t);
#pragma unroll
      for (int i = 0; i < P2; ++i) S2[i] = t[i];
    } else {
      S2[0] = 1.0;
    }
    row = ts.out_row[tb + tid];
  }
  for (int r = 0; r < nrhs; ++r) {
    __syncthreads();
    const double *src = loc + ((size_t)c * nrhs + r) * coef_stride(P);
    for (int i = tid; i < P; i += kTile) L[i] = src[i];
    __syncthreads();
    if (!active) continue;
    double v = 0.0;
#pragma unroll
    for (int i0 = 0; i0 < PO; ++i0) {
      double a1 = 0.0;
#pragma unroll
      for (int i1 = 0; i1 < P1; ++i1) {
        const double *Lp = L + (i0 * P1 + i1) * P2;
        double a2;
        if (DIM > 2) {
          a2 = 0.0;
#pragma unroll
          for (int i2 = 0; i2 < P2; ++i2) a2 += S2[i2] * Lp[i2];
        } else {
          a2 = Lp[0];
        }
        a1 += (DIM > 1 ? S1[i1] : 1.0) * a2;
      }
      v += S0[i0] * a1;
      asm volatile("" ::: "memory");  // bound the load look-ahead to one i0 slab
    }
    out[row * nrhs + r] += v;
  }
}

#define FB_ORDER_SWITCH(CALL)                       \
  if (dim == 3) {                                   \
    switch (p) {                                    \
      case 4: CALL(4, 3); return true;              \
      case 5: CALL(5, 3); return true;              \
      case 6: CALL(6, 3); return true;              \
      case 7: CALL(7, 3); return true;              \
      case 8: CALL(8, 3); return true;              \
      case 9: CALL(9, 3); return true;              \
      case 11: CALL(11, 3); return true;            \
      default: return false;                        \
    }                                               \
  } else if (dim == 2) {                            \
    switch (p) {                                    \
      case 6: CALL(6, 2); return true;              \
      case 7: CALL(7, 2); return true;              \
      case 9: CALL(9, 2); return true;              \
      case 11: CALL(11, 2); return true;            \
      default: return false;                        \
    }                                               \
  }                                                 \
  return false;

bool launch_p2m_fast(int n_leaves, const int *leaves, const int *ptb, const int *pte, const double *sx, const double *sy,
                     const double *sz, const double *w, size_t n, const double *ccx, const double *ccy,
                     const double *ccz, const double *chalf, const double *tnodes, int p, int dim, int nrhs,
                     double *mult, cudaStream_t s) {
  static const bool off = [] {
    const char *v = std::getenv("FB_GENERIC_TRANSFERS");
    return v && v[0] == '1';
  }();
  if (off) return false;
#define CALL(PO, D) \
  FB_LAUNCH((k_p2m_t<PO, D>), n_leaves, 256, 0, s, leaves, ptb, pte, sx, sy, sz, w, n, ccx, ccy, ccz, chalf, tnodes, nrhs, mult)
  FB_ORDER_SWITCH(CALL)
#undef CALL
}

bool launch_l2p_fast(const TargetSet &ts, const int *leaf_cell, const double *loc, const double *ccx, const double *ccy,
                     const double *ccz, const double *chalf, const double *tnodes, int p, int dim, int nrhs,
                     double *out, cudaStream_t s) {
  static const bool off = [] {
    const char *v = std::getenv("FB_GENERIC_TRANSFERS");
    return v && v[0] == '1';
  }();
  if (off) return false;
#define CALL(PO, D) \
  FB_LAUNCH((k_l2p_t<PO, D>), ts.max_tiles, kTile, 0, s, ts, leaf_cell, loc, ccx, ccy, ccz, chalf, tnodes, nrhs, out)
  FB_ORDER_SWITCH(CALL)
#undef CALL
}

}  // namespace fb

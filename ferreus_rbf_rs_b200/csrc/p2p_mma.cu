// EXPERIMENT (opt-in, fb_set_sqrt_mode(3)): P2P with the squared distances on the FP64 tensor cores (values only; linear /
// cubic / spheroidal families).  Reference: particle_to_particle bbfmm.rs:1162-1251, distance_sq utils.rs:230-237.
//
// Idea: the FMA-pipe P2P body (direct.cu, k_leaf_warp) is 10 FP64 instructions per pair, 6 of them the squared distance.
// With coordinates taken relative to a point of the target warp,
//     r^2 = |t|^2 + |s|^2 - 2 t.s = C + A B,   A = [tx ty tz |t|^2] (8 targets x 4),  B = [-2sx -2sy -2sz 1]^T (4 x 8 sources),
// one mma.sync.m8n8k4.f64 with C = |s|^2 yields 64 squared distances and leaves 4 FMA-pipe instructions per pair.
//
// Result on B200 (tools/dmma_bench.cu "mix", profiles/r1_dmma_dfma_mix.txt): DMMA and DFMA do NOT overlap — 4 DMMA +
// 32 DFMA per trip take the SUM of the two times (2.40 ms against 1.08 + 1.09 ms): the FP64 tensor instruction runs on
// the same datapath as DFMA (both peak at 37 TFLOP/s).  One m8n8k4 costs the pipe 8 DFMA-equivalents per 64 pairs
// against 12 for the explicit differences, so the kernel only gains 4 % (3.86 -> 3.71 ms at 1M points) and is NOT the
// default; the symmetric FMA-pipe kernel (p2p_sym.cu) halves the evaluations instead.  Kept as the measured record of
// that experiment and for its parity test (tests/test_gpu_fmm.py::test_p2p_mma_matches_fma_path).
//
// Cancellation: the expansion loses <= 9 eps max(|t|^2, |s|^2) absolutely.  |s|^2 is staged with a bias of 2^-48 of that
// maximum so every expanded r^2 stays positive, and any pair whose r^2 falls under 2^-14 of the maximum (tile-wide) is
// fixed up on a divergent cold path from coordinate differences, as the reference computes it (the diagonal 8 x 8 blocks,
// otherwise ~1e-4 of the blocks for uniform points).  Pairs that stay on the tensor-core path carry a relative error in r
// below 2^-35 (3e-11) at the threshold, falling as 1/r^2.
#include "fmm.h"

#include <cstdlib>

namespace fb {

__device__ __forceinline__ void dmma884_abc(double &d0, double &d1, double a, double b, double c0, double c1) {
  // fragments: A[lane>>2][lane&3], B[lane&3][lane>>2], C/D[lane>>2][(lane&3)*2 + {0,1}]
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

constexpr int kMmaWPC = 4;           // warps per CTA
constexpr int kMmaThrShift = 14;     // fix up when r^2 < 2^-14 max(|t|^2, |s|^2)
constexpr int kMmaBiasShift = 48;    // |s|^2 staged with + 2^-48 max(|t|^2, |s|^2): above the 9 eps max error bound of the expansion

template <int NR>
struct MmaWarpSmem {
  double b[2][32 * 4];   // per source {-2x, -2y, -2z, 1}: the B fragment of column group c is b[32 c + lane]
  double ss[2][32];      // |s|^2 (+ bias): the C fragment of column group c is ss[8 c + 2 (lane & 3) + {0, 1}]
  double w[2][NR][32];   // staged weights (x kernel_weight_scale)
  double t[4][32];       // targets {x, y, z, |t|^2} relative to the warp's origin
};

// One staged source tile against NG groups of 8 targets, in units of (8-source column group, <= 2 target groups):
// the tensor-core products of a unit are issued one unit ahead of their use, so the DMMA latency hides under the
// previous unit's arithmetic; a unit then runs its 2 G kernel evaluations per lane as independent chains and ends with
// one merged threshold test that guards the cold fix-up path.
template <int NR, int G, int G0>
__device__ __forceinline__ void mma_issue(double (&d)[2][2], const MmaWarpSmem<NR> &sm, const int buf, const int c,
                                          const double (&af)[4], const int lane) {
  const double bf = sm.b[buf][32 * c + lane];
  const double2 cs = *reinterpret_cast<const double2 *>(&sm.ss[buf][8 * c + 2 * (lane & 3)]);
#pragma unroll
  for (int g = 0; g < G; ++g) dmma884_abc(d[g][0], d[g][1], af[G0 + g], bf, cs.x, cs.y);
}

template <int FAM, int NR, int G, int G0>
__device__ __forceinline__ void mma_consume(const double (&d)[2][2], const MmaWarpSmem<NR> &sm, const int buf,
                                            const int c, const int thr, double (&acc)[4][NR], const int lane,
                                            const KParams &kp) {
  const int row = lane >> 2, kk = lane & 3;
  double w0[NR], w1[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const double2 w2 = *reinterpret_cast<const double2 *>(&sm.w[buf][r][8 * c + 2 * kk]);
    w0[r] = w2.x;
    w1[r] = w2.y;
  }
  // the staged |s|^2 carries a positive bias above the rounding error of the expansion: d > 0 always, so the
  // zero-distance guard of the square-root seed is not needed here
  double v[G][2];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    v[g][0] = kernel_mag<FAM, true, false>(d[g][0], kp);
    v[g][1] = kernel_mag<FAM, true, false>(d[g][1], kp);
  }
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      kernel_acc<FAM>(acc[G0 + g][r], v[g][0], w0[r]);
      kernel_acc<FAM>(acc[G0 + g][r], v[g][1], w1[r]);
    }
  int lo = min(__double2hiint(d[0][0]), __double2hiint(d[0][1]));
#pragma unroll
  for (int g = 1; g < G; ++g) lo = min(lo, min(__double2hiint(d[g][0]), __double2hiint(d[g][1])));
  if (lo < thr) {
    // cold fix-up: replace the tensor-core value of each flagged pair by the one from coordinate differences
    // (x, y, z in order, utils.rs:230-237): acc += (k(exact) - k(expansion)) w
    const double *bt = sm.b[buf];
    const int S = 8 * c + 2 * kk;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int T = 8 * (G0 + g) + row;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (__double2hiint(d[g][u]) < thr) {
          const double dx = sm.t[0][T] + 0.5 * bt[4 * (S + u) + 0], dy = sm.t[1][T] + 0.5 * bt[4 * (S + u) + 1],
                       dz = sm.t[2][T] + 0.5 * bt[4 * (S + u) + 2];
          double q = dx * dx;
          q += dy * dy;
          q += dz * dz;
          const double fix = kernel_mag<FAM, true, true>(q, kp) - v[g][u];
#pragma unroll
          for (int r = 0; r < NR; ++r) kernel_acc<FAM>(acc[G0 + g][r], fix, u == 0 ? w0[r] : w1[r]);
        }
      }
    }
  }
}

template <int FAM, int NR, int NG>
__device__ __forceinline__ void mma_tile(const MmaWarpSmem<NR> &sm, const int buf, const int m, const int thr,
                                         const double (&af)[4], double (&acc)[4][NR], const int lane,
                                         const KParams &kp) {
  const int ncg = (m + 7) >> 3;  // 1..4, warp-uniform
  double da[2][2], db[2][2];
  if (NG > 2) {  // two units per column group: target groups {0, 1} and {2[, 3]}
    constexpr int G1 = NG > 2 ? NG - 2 : 1;
    mma_issue<NR, 2, 0>(da, sm, buf, 0, af, lane);
    for (int c = 0; c < ncg; ++c) {
      mma_issue<NR, G1, 2>(db, sm, buf, c, af, lane);
      mma_consume<FAM, NR, 2, 0>(da, sm, buf, c, thr, acc, lane, kp);
      if (c + 1 < ncg) mma_issue<NR, 2, 0>(da, sm, buf, c + 1, af, lane);
      mma_consume<FAM, NR, G1, 2>(db, sm, buf, c, thr, acc, lane, kp);
    }
  } else {  // one unit per column group, column groups ping-pong
    constexpr int G = NG > 2 ? 2 : NG;
    mma_issue<NR, G, 0>(da, sm, buf, 0, af, lane);
    for (int c = 0; c < ncg; c += 2) {
      if (c + 1 < ncg) mma_issue<NR, G, 0>(db, sm, buf, c + 1, af, lane);
      mma_consume<FAM, NR, G, 0>(da, sm, buf, c, thr, acc, lane, kp);
      if (c + 1 < ncg) {
        if (c + 2 < ncg) mma_issue<NR, G, 0>(da, sm, buf, c + 2, af, lane);
        mma_consume<FAM, NR, G, 0>(db, sm, buf, c + 1, thr, acc, lane, kp);
      }
    }
  }
}

template <int FAM, int NR>
__global__ void __launch_bounds__(kMmaWPC * 32, NR >= 8 ? 2 : (NR >= 4 ? 3 : 4)) k_p2p_mma(const DirectArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * kMmaWPC + warp;
  const int tile = (int)(gw >> 2), sub = (int)(gw & 3);
  if (tile >= *a.ts.n_tiles_dev) return;
  const int li = a.ts.tile_leaf[tile];
  const int tb = a.ts.leaf_begin[li] + a.ts.tile_off[tile] + sub * 32;
  const int cnt = min(32, a.ts.leaf_end[li] - tb);
  if (cnt <= 0) return;
  long long e = a.u_ptr[li];
  const long long e_end = a.u_ptr[li + 1];
  if (e >= e_end) return;

  extern __shared__ __align__(16) unsigned char dsm_raw[];
  MmaWarpSmem<NR> &sm = reinterpret_cast<MmaWarpSmem<NR> *>(dsm_raw)[warp];

  // ---- targets, relative to the first target of the warp
  const double ox = a.ts.x[tb], oy = a.ts.y[tb], oz = a.ts.z[tb];
  int tt_hi;
  {
    double tx = 0, ty = 0, tz = 0;
    if (lane < cnt) {
      tx = a.ts.x[tb + lane] - ox;
      ty = a.ts.y[tb + lane] - oy;
      tz = a.ts.z[tb + lane] - oz;
    }
    double tt = tx * tx;
    tt = fma(ty, ty, tt);
    tt = fma(tz, tz, tt);
    sm.t[0][lane] = tx;
    sm.t[1][lane] = ty;
    sm.t[2][lane] = tz;
    sm.t[3][lane] = tt;
    tt_hi = __reduce_max_sync(0xffffffffu, __double2hiint(tt));
  }
  __syncwarp();
  const int row = lane >> 2, kk = lane & 3;
  const int ngroups = (cnt + 7) >> 3;
  double af[4];  // A fragments: row = target 8 g + (lane >> 2), column kk of {x, y, z, |t|^2}
#pragma unroll
  for (int g = 0; g < 4; ++g) af[g] = sm.t[kk][8 * g + row];
  double acc[4][NR];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[g][r] = 0.0;

  // ---- source tiles over the merged U ranges: registers -> shared memory, double buffered
  int rb = a.u_begin[e], rn = a.u_count[e], c0 = 0;
  double reg[4 + NR];
  auto fetch = [&](int &m) {
    m = 0;
    if (e >= e_end) return;
    m = min(32, rn - c0);
#pragma unroll
    for (int k = 0; k < 4 + NR; ++k) reg[k] = 0.0;
    if (lane < m) {
      const int s = rb + c0 + lane;
      const double x = a.sx[s] - ox, y = a.sy[s] - oy, z = a.sz[s] - oz;
      double ss = x * x;
      ss = fma(y, y, ss);
      ss = fma(z, z, ss);
      reg[0] = -2.0 * x;
      reg[1] = -2.0 * y;
      reg[2] = -2.0 * z;
      reg[3] = ss;
#pragma unroll
      for (int r = 0; r < NR; ++r) reg[4 + r] = a.w[(size_t)(a.rhs0 + r) * a.n + s] * kernel_weight_scale<FAM, true>();
    }
    c0 += 32;
    if (c0 >= rn) {
      ++e;
      c0 = 0;
      if (e < e_end) {
        rb = a.u_begin[e];
        rn = a.u_count[e];
      }
    }
  };
  auto stash = [&](int buf, int m) -> int {  // returns the fix-up threshold (high word) of the tile
    if (m <= 0) return 0;
    const int mx_hi = max(__reduce_max_sync(0xffffffffu, __double2hiint(reg[3])), tt_hi);  // max(|t|^2, |s|^2)
    // real sources: |s|^2 + bias, so every expanded r^2 stays positive; padding (zero weight): r^2 >= the maximum, never
    // flagged for the fix-up path
    const double bias = __hiloint2double(max(mx_hi - (kMmaBiasShift << 20), 0x01000000), 0);
    const double ssb = lane < m ? reg[3] + bias : __hiloint2double(max(mx_hi, 0x01000000), 0);
    *reinterpret_cast<double2 *>(&sm.b[buf][4 * lane]) = make_double2(reg[0], reg[1]);
    *reinterpret_cast<double2 *>(&sm.b[buf][4 * lane + 2]) = make_double2(reg[2], 1.0);
    sm.ss[buf][lane] = ssb;
#pragma unroll
    for (int r = 0; r < NR; ++r) sm.w[buf][r][lane] = reg[4 + r];
    return max(mx_hi - (kMmaThrShift << 20), 0);
  };

  int m_cur = 0, m_next = 0, buf = 0;
  fetch(m_cur);
  int thr = stash(0, m_cur);
  __syncwarp();
  while (m_cur > 0) {
    fetch(m_next);  // global loads of the next tile overlap the arithmetic below
    switch (ngroups) {
      case 1: mma_tile<FAM, NR, 1>(sm, buf, m_cur, thr, af, acc, lane, a.kp); break;
      case 2: mma_tile<FAM, NR, 2>(sm, buf, m_cur, thr, af, acc, lane, a.kp); break;
      case 3: mma_tile<FAM, NR, 3>(sm, buf, m_cur, thr, af, acc, lane, a.kp); break;
      default: mma_tile<FAM, NR, 4>(sm, buf, m_cur, thr, af, acc, lane, a.kp); break;
    }
    buf ^= 1;
    const int thr_next = stash(buf, m_next);
    __syncwarp();
    m_cur = m_next;
    thr = thr_next;
  }

  // ---- quad reduction (the 4 lanes of a row hold its 8 columns), lane kk == 0 writes the row
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double v = acc[g][r];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      acc[g][r] = v;
    }
  if (kk == 0) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int T = 8 * g + row;
      if (T < cnt) {
        const size_t orow = a.ts.out_row[tb + T];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          double *o = a.out + orow * a.nrhs + a.rhs0 + r;
          if (a.atomic_out) atomicAdd(o, acc[g][r]);
          else *o += acc[g][r];
        }
      }
    }
  }
}

template <int FAM, int NR>
static void launch_p2p_mma_nr(const DirectArgs &a, cudaStream_t s) {
  const size_t smem = sizeof(MmaWarpSmem<NR>) * kMmaWPC;
  const unsigned grid = (unsigned)(((long long)a.ts.max_tiles * 4 + kMmaWPC - 1) / kMmaWPC);
  if (smem > 48 * 1024)
    FB_CUDA(cudaFuncSetAttribute(k_p2p_mma<FAM, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  FB_LAUNCH((k_p2p_mma<FAM, NR>), grid, kMmaWPC * 32, smem, s, a);
}

template <int FAM>
static void p2p_mma_fam(DirectArgs a, cudaStream_t s) {
  int r = 0;
  while (r < a.nrhs) {
    a.rhs0 = r;
    const int left = a.nrhs - r;
    if (left >= 8) {
      launch_p2p_mma_nr<FAM, 8>(a, s);
      r += 8;
    } else if (left >= 4) {
      launch_p2p_mma_nr<FAM, 4>(a, s);
      r += 4;
    } else if (left >= 2) {
      launch_p2p_mma_nr<FAM, 2>(a, s);
      r += 2;
    } else {
      launch_p2p_mma_nr<FAM, 1>(a, s);
      r += 1;
    }
  }
}

// true when the tensor-core P2P serves this call (values only, mode 3, a family with a FAST hot loop)
bool p2p_mma_applicable(const DirectArgs &a) {
  if (a.gout || a.kp.fast != 3 || a.ts.max_tiles <= 0) return false;
  return a.kp.fam == KF_LINEAR || a.kp.fam == KF_CUBIC || a.kp.fam == KF_SPH;
}

void launch_p2p_mma(const DirectArgs &a, cudaStream_t s) {
  switch (a.kp.fam) {
    case KF_LINEAR: p2p_mma_fam<KF_LINEAR>(a, s); break;
    case KF_CUBIC: p2p_mma_fam<KF_CUBIC>(a, s); break;
    default: p2p_mma_fam<KF_SPH>(a, s); break;
  }
}

}  // namespace fb

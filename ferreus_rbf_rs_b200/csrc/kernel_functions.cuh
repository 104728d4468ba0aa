// The closed kernel registry of ferreus_rbf_utils (utils.rs:558-571) as host+device functors.
// Math follows rbf_kernels.rs:25-317, non_rbf_kernels.rs:20-163, constants.rs:13-50.
// Kernels are selected by template parameter (no function pointers on the device).
#pragma once
#include <cmath>

#include "../../include/ferreus_b200.h"

#ifdef __CUDACC__
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD inline
#endif

namespace fb {

// kernel families (the four spheroidal orders share one family with a runtime exponent)
enum KFam { KF_LINEAR = 0, KF_TPS, KF_CUBIC, KF_SPH, KF_LAPLACE, KF_R2, KF_R4, KF_COUNT };

struct KParams {
  int fam;
  int fast;           // 1: second-order square roots in the direct-sum hot loops (see kernel_mag)
  int pw;             // spheroidal POW (1..4), rbf_kernels.rs:176-203
  double s2;          // s^2, s = range_scaling / base_range
  double ip2;         // inflexion_point^2
  double near_slope;  // total_sill * linear_slope * s
  double far_coef;    // total_sill * inv_y_intercept
  double total_sill;
};

inline bool make_kparams(const fb_kernel_params &k, KParams &out) {
  static const double C[4][4] = {// inflexion_point, linear_slope, range_scaling, inv_y_intercept (constants.rs:21-50)
                                 {0.5000000000, 0.7500000000, 2.6798340586, 0.8734640537},
                                 {0.4082482905, 1.0206207262, 1.5822795750, 0.8575980168},
                                 {0.3535533906, 1.2374368671, 1.2008676644, 0.8494862533},
                                 {0.3162277660, 1.4230249471, 1.0000000000, 0.8445585690}};
  out = KParams{};
  out.total_sill = k.total_sill;
  switch (k.kernel_type) {
    case FB_KERNEL_LINEAR: out.fam = KF_LINEAR; return true;
    case FB_KERNEL_THIN_PLATE_SPLINE: out.fam = KF_TPS; return true;
    case FB_KERNEL_CUBIC: out.fam = KF_CUBIC; return true;
    case FB_KERNEL_LAPLACIAN: out.fam = KF_LAPLACE; return true;
    case FB_KERNEL_ONE_OVER_R2: out.fam = KF_R2; return true;
    case FB_KERNEL_ONE_OVER_R4: out.fam = KF_R4; return true;
    case FB_KERNEL_SPHEROIDAL3:
    case FB_KERNEL_SPHEROIDAL5:
    case FB_KERNEL_SPHEROIDAL7:
    case FB_KERNEL_SPHEROIDAL9: {
      const int o = k.kernel_type - FB_KERNEL_SPHEROIDAL3;
      const double s = C[o][2] / k.base_range;  // rbf_kernels.rs:229-238
      out.fam = KF_SPH;
      out.pw = o + 1;
      out.s2 = s * s;
      out.ip2 = C[o][0] * C[o][0];
      out.near_slope = k.total_sill * C[o][1] * s;
      out.far_coef = k.total_sill * C[o][3];
      return true;
    }
    default: return false;
  }
}

constexpr double kEps = 2.220446049250313e-16;  // f64::EPSILON

// value from squared distance
template <int FAM>
FB_HD double kernel_value(double r2, const KParams &kp) {
  if (FAM == KF_LINEAR) {
    return -sqrt(r2);
  } else if (FAM == KF_TPS) {  // r^2 ln r, 0 when r < eps (rbf_kernels.rs:77-83)
    const double r = sqrt(r2);
    return (r < kEps) ? 0.0 : (r * r) * log(r);
  } else if (FAM == KF_CUBIC) {
    const double r = sqrt(r2);
    return r * r * r;
  } else if (FAM == KF_SPH) {  // rbf_kernels.rs:243-256
    const double sr2 = kp.s2 * r2;
    if (sr2 <= kp.ip2) return kp.total_sill - kp.near_slope * sqrt(r2);
    const double t = 1.0 + sr2;
    const double t2 = t * t;  // t.powi(POW): binary exponentiation, as LLVM lowers powi
    const double tp = kp.pw == 1 ? t : (kp.pw == 2 ? t2 : (kp.pw == 3 ? t2 * t : t2 * t2));
    return kp.far_coef / (tp * sqrt(t));
  } else if (FAM == KF_LAPLACE) {
    const double r = sqrt(r2);
    return (r < kEps) ? 0.0 : 1.0 / r;
  } else if (FAM == KF_R2) {
    const double r = sqrt(r2);
    return (r < kEps) ? 0.0 : 1.0 / (r * r);
  } else {
    const double r = sqrt(r2);
    const double rr = r * r;
    return (r < kEps) ? 0.0 : 1.0 / (rr * rr);
  }
}

// value and gradient factor: grad = fac * (target - source); zero gradient when r2 <= eps
template <int FAM>
FB_HD void kernel_value_grad(double r2, const KParams &kp, double &val, double &fac) {
  const bool small = r2 <= kEps;
  const double r = sqrt(r2);
  if (FAM == KF_LINEAR) {
    val = -r;
    fac = small ? 0.0 : -1.0 / r;
  } else if (FAM == KF_TPS) {
    const double lr = small ? 0.0 : log(r);
    val = small ? 0.0 : r2 * lr;
    fac = small ? 0.0 : 2.0 * lr + 1.0;
  } else if (FAM == KF_CUBIC) {
    val = small ? 0.0 : r2 * r;
    fac = small ? 0.0 : 3.0 * r;
  } else if (FAM == KF_SPH) {
    val = kernel_value<KF_SPH>(r2, kp);
    const double sr2 = kp.s2 * r2;
    if (small) {
      fac = 0.0;
    } else if (sr2 <= kp.ip2) {
      fac = -kp.near_slope * (1.0 / r);
    } else {
      const double t = 1.0 + sr2;
      const double p = (double)kp.pw + 0.5;
      fac = -2.0 * p * kp.s2 * kp.far_coef / pow(t, p + 1.0);
    }
  } else if (FAM == KF_LAPLACE) {
    const double ir = small ? 0.0 : 1.0 / r;
    val = ir;
    fac = -(ir * ir * ir);
  } else if (FAM == KF_R2) {
    val = small ? 0.0 : 1.0 / r2;
    fac = small ? 0.0 : -2.0 * (1.0 / (r2 * r2));
  } else {
    val = small ? 0.0 : 1.0 / (r2 * r2);
    fac = small ? 0.0 : -4.0 * (1.0 / (r2 * r2 * r2));
  }
}


#ifdef __CUDACC__
// ---- device fast path ---------------------------------------------------------------------------
// sqrt(a) and 1/sqrt(a) for a > 0 (normal): MUFU.RSQ64H seed y0 (~2^-21 relative: it reads the high word only)
// refined with one third-order step — full double precision (~1 ulp).  tools/fp64_ubench.cu measures the rates this
// is designed around: DFMA 1.90 warp-instr/clk/SM, MUFU.RSQ64H 0.50, IEEE sqrt() 2.6x slower than this sequence.
__device__ __forceinline__ double rsq64h(double a) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
  return y0;
}
// With y0 = (1 + d) / sqrt(a):  r = a y0,  e = 1 - r y0 = 1 - (1 + d)^2,  and  (1 - e)^(-1/2) = 1 + e/2 + 3e^2/8 + O(e^3):
//   sqrt(a)   = r  + (r  e)(1/2 + 3e/8),      1/sqrt(a) = y0 + (y0 e)(1/2 + 3e/8)        (remainder 5e^3/16 < 2^-61)
// 5 FP64 operations, branch-free, no IEEE slow path, no exponent fix-up of the seed (hot_sqrt / hot_rsqrt below).
// seed taken from max(a, 2^-1022) (one integer max on the high word): a == 0 then gives r = 0 * finite = 0, e = 1 and
// the results sqrt -> exactly 0, a * rsqrt -> exactly 0, without a select on the FP64 result
__device__ __forceinline__ double seed_operand(double a) {
  return __hiloint2double(max(__double2hiint(a), 0x00100000), __double2loint(a));
}
// ---- hot-loop variants -------------------------------------------------------------------------------
// FAST = false: third-order step, ~1 ulp (measured 1.7e-16 max relative error, tools/fp64_ubench.cu).
// FAST = true : second-order step r (3 - r y0) / 2: 3 FP64 operations instead of 5, relative error
//               -1.5 d^2 with |d| <= 2^-20.1 the seed error: measured <= 1.24e-12.  Default; fb_set_sqrt_mode(0) or
//               FB_SQRT=exact selects the third-order step.  The parity gates (1e-10 matvec, 1e-8 interpolant) hold
//               in both modes (tests/test_gpu_fmm.py::test_sqrt_modes).
__device__ __forceinline__ double half_of(double y) {  // y / 2 as an exponent decrement on the integer pipe
  return __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
}
// FAST returns 2 sqrt(a) = r (3 - r y0): the factor 1/2 rides on the weights (kernel_weight_scale), which saves the
// exponent fix-up of the seed — one instruction per pair in loops that are issue- as much as FP64-bound
// (tools/fp64_ubench.cu: far body 2.40 -> 2.58 Tpair/s at 16 warps/SM)
template <bool FAST, bool ZERO_OK>
__device__ __forceinline__ double hot_sqrt(double a) {  // ZERO_OK: a == 0 -> exactly 0 (seed from max(a, 2^-1022))
  const double y0 = rsq64h(ZERO_OK ? seed_operand(a) : a);
  const double r = a * y0;
  if (FAST) return r * fma(-r, y0, 3.0);
  const double e = fma(-r, y0, 1.0);
  const double c = fma(e, 0.375, 0.5);
  return fma(r * e, c, r);
}
template <bool FAST, bool ZERO_OK>
__device__ __forceinline__ double hot_rsqrt(double a) {  // ZERO_OK: finite for a == 0, a * result == 0
  const double y0 = rsq64h(ZERO_OK ? seed_operand(a) : a);
  if (FAST) {
    const double e = fma(-(a * y0), half_of(y0), 0.5);
    return fma(y0, e, y0);
  }
  const double e = fma(-(a * y0), y0, 1.0);
  const double c = fma(e, 0.375, 0.5);
  return fma(y0 * e, c, y0);
}

// Sign-stripped kernel value + signed accumulate: for the linear kernel (-r) the negation rides on the DFMA
// operand modifier instead of costing an FP64 instruction per pair (the compiler hoists a plain negation above a
// select).  NEAR = true: r2 may be 0 (P2P, self term included); false: well-separated cells (M2P, P2L), r2 > 0.
template <int FAM, bool FAST, bool NEAR>
__device__ __forceinline__ double kernel_mag(double r2, const KParams &kp) {
  if (FAM == KF_LINEAR) {
    return hot_sqrt<FAST, NEAR>(r2);
  } else if (FAM == KF_CUBIC) {
    return r2 * hot_sqrt<FAST, NEAR>(r2);
  } else if (FAM == KF_SPH) {
    const double sr2 = kp.s2 * r2;
    const bool near = sr2 <= kp.ip2;
    const double t = 1.0 + sr2;
    const double y = hot_rsqrt<FAST, NEAR>(near ? r2 : t);
    const double vn = kp.total_sill - kp.near_slope * (r2 * y);
    const double y2 = y * y;
    double tp = y2;
    for (int i = 1; i < kp.pw; ++i) tp *= y2;  // t^-pw
    const double vf = kp.far_coef * (y * tp);
    return near ? vn : vf;
  } else if (FAM == KF_TPS) {  // r^2 ln r = r2 * ln(r2) / 2
    const double v = (0.5 * r2) * log(r2);
    return (!NEAR || r2 >= kEps * kEps) ? v : 0.0;
  } else {
    const double y = hot_rsqrt<false, false>(r2);
    const double y2 = y * y;
    const double v = FAM == KF_LAPLACE ? y : (FAM == KF_R2 ? y2 : y2 * y2);
    return (!NEAR || r2 >= kEps * kEps) ? v : 0.0;
  }
}
template <int FAM>
__device__ __forceinline__ void kernel_acc(double &acc, double mag, double w) {
  if (FAM == KF_LINEAR) acc -= mag * w;
  else acc += mag * w;
}
// kernel_mag<FAM, FAST, .> returns value / kernel_weight_scale<FAM, FAST>(): the callers stage weights (and the
// multipoles of the M2P half) multiplied by this scale
template <int FAM, bool FAST>
__host__ __device__ constexpr double kernel_weight_scale() {
  return (FAST && (FAM == KF_LINEAR || FAM == KF_CUBIC)) ? 0.5 : 1.0;
}
// families whose hot loop has a FAST variant (the others ignore the switch)
template <int FAM>
constexpr bool kernel_has_fast() {
  return FAM == KF_LINEAR || FAM == KF_CUBIC || FAM == KF_SPH;
}
#endif

// runtime-dispatched host evaluation (operator precompute, dense domain matrices)
inline double kernel_value_rt(double r2, const KParams &kp) {
  switch (kp.fam) {
    case KF_LINEAR: return kernel_value<KF_LINEAR>(r2, kp);
    case KF_TPS: return kernel_value<KF_TPS>(r2, kp);
    case KF_CUBIC: return kernel_value<KF_CUBIC>(r2, kp);
    case KF_SPH: return kernel_value<KF_SPH>(r2, kp);
    case KF_LAPLACE: return kernel_value<KF_LAPLACE>(r2, kp);
    case KF_R2: return kernel_value<KF_R2>(r2, kp);
    default: return kernel_value<KF_R4>(r2, kp);
  }
}

}  // namespace fb

// Host-side setup of the RBF solve: duplicate removal, monomial bases, DDM hierarchy, per-domain special
// points / Lagrange Q matrices.  Reference semantics: ferreus_rbf/src/rbf.rs:1391-1467, polynomials.rs:15-130,
// common.rs:246-320, preconditioning/domain_decomposition.rs:67-346, domain.rs:153-383 (host part).
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../include/ferreus_rbf_b200.h"
#include "kernel_functions.cuh"

namespace fb {

struct Settings {  // InterpolantSettings after set_basis_size (interpolant_config.rs:229-264)
  int kernel_type = 0, drift = 1, spheroidal_order = 3;
  double nugget = 0, base_range = 1, total_sill = 1, tolerance = 1e-6;
  int tolerance_type = 0;
  int basis_size = 0, polynomial_degree = -1;
  fb_kernel_params kparams{};
};

bool resolve_settings(const fr_settings &in, int dim, Settings &out, std::string &err);

void cheb_cube_scaling(const double *pts, const int64_t *idx, size_t n, int dim, double *translation,
                       double *scale);                                         // common.rs:299-320
// monomials (n x basis, row-major) of points pts[idx[i]] (idx may be null), polynomials.rs:15-62
void evaluate_monomials(const double *pts, const int64_t *idx, size_t n, int dim, int degree, int basis,
                        const double *translation, const double *scale, double *out);
std::vector<int64_t> remove_duplicates(const double *pts, size_t n, int dim, const KParams &kp);  // rbf.rs:1418-1467
double duplicate_cutoff_distance(double h_ref, const KParams &kp);                                // rbf.rs:1391-1416
std::vector<int> farthest_point_sampling(const double *pts, const int64_t *idx, size_t n, int dim, size_t wanted,
                                         size_t seed);                                            // common.rs:246-287

struct DomainHost {
  std::vector<int64_t> idx;      // overlapping_point_indices (special points first after prepare())
  std::vector<uint8_t> mask;     // internal_points_mask aligned with idx
  std::vector<double> extents;   // [mins..., maxs...]
  int rank = 0;                  // number of special points (0 without polynomials)
  std::vector<double> qtop;      // rank x (n - rank), row-major  (Q_top = -Lagrange(non-special)^T)
  std::vector<double> sp_inv;    // rank x rank inverse of the special-point monomials (poly recovery)
  bool solve_for_poly = false;
  // host part of Domain::factorise (domain.rs:163-330): unisolvent columns, special points, reorder, Q_top
  // mono_pts: coordinates the monomials are evaluated at (global trend: the inverse-transformed points, domain.rs:169-175)
  void prepare(const double *pts, int dim, const Settings &s, bool solve_for_poly_, const double *mono_pts = nullptr);
};

struct LevelHost {
  std::vector<int64_t> point_indices;
  std::vector<DomainHost> domains;
};

// DDMTree::new without the factorisations (domain_decomposition.rs:67-346)
// on_level(l, level, is_coarse) is called as soon as level l is complete (its domains prepared), before the next level
// is built
using LevelCallback = std::function<void(size_t, const LevelHost &, bool)>;
std::vector<LevelHost> build_ddm(const double *pts, size_t n, int dim, const Settings &s, const fr_params &p,
                                 const double *mono_pts = nullptr, const LevelCallback *on_level = nullptr);

// thin Q of an n x m matrix (row-major in, row-major out), rbf.rs:493-495
void thin_q_rowmajor(const double *a, size_t n, int m, double *q);

}  // namespace fb

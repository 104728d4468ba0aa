// Host-side precompute of the Chebyshev interpolation / M2M / compressed M2L operators.
// Reference semantics: ferreus_bbfmm/src/chebyshev.rs:32-814, aca.rs:23-247.
// Dense factorizations (faer 0.23.2 in the reference: thin QR aca.rs:175-181, SVD aca.rs:186,
// chebyshev.rs:763) are restated here as Householder QR and one-sided Jacobi SVD.
#pragma once
#include <cstdint>
#include <vector>

#include "kernel_functions.cuh"

namespace fb {

// column-major dense matrix
struct Mat {
  int rows = 0, cols = 0;
  std::vector<double> a;
  Mat() = default;
  Mat(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
  double &operator()(int i, int j) { return a[(size_t)j * rows + i]; }
  double operator()(int i, int j) const { return a[(size_t)j * rows + i]; }
};

void thin_qr(const Mat &A, Mat &Q, Mat &R);                              // A = Q R, Q m x n, R n x n
void jacobi_svd(const Mat &A, Mat &U, std::vector<double> &S, Mat &V);   // A = U diag(S) V^T, S descending
int singular_values_cutoff(const std::vector<double> &sigma, double eps);  // aca.rs:210-247

struct M2LOperator {
  int rank = 0;  // P when uncompressed
  Mat U;         // P x rank   (uncompressed: the dense K, P x P)
  Mat Vt;        // rank x P   (empty when uncompressed)
};

struct Operators {
  int p = 0, dim = 0, P = 0, n_ref = 0, n_perm = 0, n_vec = 0;
  std::vector<double> nodes;      // p, ascending (chebyshev.rs:32-40)
  std::vector<double> tnodes;     // p x p row-major: T_k(node_m) at [m*p + k]
  std::vector<double> child_s;    // 2 x p x p: [h][i][m] = S_m((node_i -/+ 1)/2)  (chebyshev.rs:146-180)
  std::vector<int32_t> perm;      // n_perm x P  permutation_indices
  std::vector<int32_t> inv_perm;  // n_perm x P  inverse_permutations
  std::vector<int32_t> perm_lookup;  // 7^dim
  std::vector<int32_t> ref_lookup;   // 7^dim
  std::vector<int32_t> ref_vecs;     // n_ref x dim
  // [level-2][ref]
  std::vector<std::vector<M2LOperator>> m2l;

  void build(int order, int dim_, double radius, int depth, const KParams &kp, int compression, double eps);
  // build() behind a small process-wide cache keyed by every argument (bit patterns of the doubles): the trees a model
  // builds over one point set — the solver's, the evaluator's, one per one-shot evaluate — ask for identical operators,
  // and at order 11 the ACA + recompression of the 48 of them is 0.26 s per tree.  Returns true on a cache hit.
  bool build_cached(int order, int dim_, double radius, int depth, const KParams &kp, int compression, double eps);
  size_t bytes() const;
};

}  // namespace fb

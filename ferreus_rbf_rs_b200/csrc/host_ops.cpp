// Host-side operator precompute. See host_ops.h for the reference mapping.
#include "host_ops.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace fb {

// ------------------------------------------------------------------------------------ dense LA
void thin_qr(const Mat &A, Mat &Q, Mat &R) {
  const int m = A.rows, n = A.cols;
  Mat W = A;
  std::vector<double> tau(n, 0.0);
  for (int k = 0; k < n; ++k) {
    double nrm = 0;
    for (int i = k; i < m; ++i) nrm += W(i, k) * W(i, k);
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) continue;
    const double alpha = W(k, k) > 0 ? -nrm : nrm;
    const double v0 = W(k, k) - alpha;
    for (int i = k + 1; i < m; ++i) W(i, k) /= v0;  // v = [1, w/v0]
    tau[k] = -v0 / alpha;                           // H = I - tau v v^T
    W(k, k) = alpha;
    for (int j = k + 1; j < n; ++j) {
      double s = W(k, j);
      for (int i = k + 1; i < m; ++i) s += W(i, k) * W(i, j);
      s *= tau[k];
      W(k, j) -= s;
      for (int i = k + 1; i < m; ++i) W(i, j) -= s * W(i, k);
    }
  }
  R = Mat(n, n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i <= j && i < m; ++i) R(i, j) = W(i, j);
  Q = Mat(m, n);
  for (int j = 0; j < n && j < m; ++j) Q(j, j) = 1.0;
  for (int k = std::min(n, m) - 1; k >= 0; --k) {
    if (tau[k] == 0.0) continue;
    for (int j = 0; j < n; ++j) {
      double s = Q(k, j);
      for (int i = k + 1; i < m; ++i) s += W(i, k) * Q(i, j);
      s *= tau[k];
      Q(k, j) -= s;
      for (int i = k + 1; i < m; ++i) Q(i, j) -= s * W(i, k);
    }
  }
}

void jacobi_svd(const Mat &A, Mat &U, std::vector<double> &S, Mat &V) {
  const int m = A.rows, n = A.cols;
  Mat W = A;
  Mat Vv(n, n);
  for (int j = 0; j < n; ++j) Vv(j, j) = 1.0;
  const double tol = 1e-15;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        const double *wp = &W.a[(size_t)p * m], *wq = &W.a[(size_t)q * m];
        for (int i = 0; i < m; ++i) {
          alpha += wp[i] * wp[i];
          beta += wq[i] * wq[i];
          gamma += wp[i] * wq[i];
        }
        if (gamma == 0.0 || std::fabs(gamma) <= tol * std::sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        double *wpm = &W.a[(size_t)p * m], *wqm = &W.a[(size_t)q * m];
        for (int i = 0; i < m; ++i) {
          const double a = wpm[i], b = wqm[i];
          wpm[i] = c * a - s * b;
          wqm[i] = s * a + c * b;
        }
        double *vp = &Vv.a[(size_t)p * n], *vq = &Vv.a[(size_t)q * n];
        for (int i = 0; i < n; ++i) {
          const double a = vp[i], b = vq[i];
          vp[i] = c * a - s * b;
          vq[i] = s * a + c * b;
        }
      }
    if (!rotated) break;
  }
  std::vector<double> sv(n);
  for (int j = 0; j < n; ++j) {
    double nrm = 0;
    for (int i = 0; i < m; ++i) nrm += W(i, j) * W(i, j);
    sv[j] = std::sqrt(nrm);
  }
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sv[a] > sv[b]; });
  U = Mat(m, n);
  V = Mat(n, n);
  S.assign(n, 0.0);
  for (int jj = 0; jj < n; ++jj) {
    const int j = order[jj];
    S[jj] = sv[j];
    const double inv = sv[j] > 0 ? 1.0 / sv[j] : 0.0;
    for (int i = 0; i < m; ++i) U(i, jj) = W(i, j) * inv;
    for (int i = 0; i < n; ++i) V(i, jj) = Vv(i, j);
  }
}

int singular_values_cutoff(const std::vector<double> &sigma, double eps) {
  const int n = (int)sigma.size();
  std::vector<double> cum(n, 0.0);
  double acc = 0;
  for (int i = n - 1; i >= 0; --i) {
    acc += sigma[i] * sigma[i];
    cum[i] = acc;
  }
  if (n == 0) return 0;
  const double eps_qr = cum[0] * eps * eps;
  for (int i = 0; i < n; ++i)
    if (cum[i] < eps_qr) return i;
  return n;
}

// ------------------------------------------------------------------------------------ Chebyshev
static void cheb_T(int p, double x, double *T) {  // chebyshev.rs:47-110
  T[0] = 1.0;
  if (p > 1) T[1] = x;
  for (int k = 2; k < p; ++k) T[k] = 2.0 * x * T[k - 1] - T[k - 2];
}

namespace {

struct AcaResult {
  Mat u, v;  // P x k each
};

// K[i, j] = k(src_i, tgt_j): rows = nodes of the offset (receiving) cell, cols = nodes of the origin cell
struct KGen {
  const std::vector<double> *src, *tgt;  // P x dim row-major
  int P, dim;
  const KParams *kp;
  double at(int i, int j) const {
    double r2 = 0;
    for (int d = 0; d < dim; ++d) {
      const double df = (*src)[(size_t)i * dim + d] - (*tgt)[(size_t)j * dim + d];
      r2 += df * df;
    }
    return kernel_value_rt(r2, *kp);
  }
};

int argmax_masked(const std::vector<double> &data, const std::vector<uint8_t> &mask) {  // aca.rs:146-161
  int best = 0;
  double best_val = 0.0;
  for (size_t i = 0; i < data.size(); ++i) {
    const double w = std::fabs(data[i]) * (double)mask[i];
    if (w > best_val) {
      best_val = w;
      best = (int)i;
    }
  }
  return best;
}

AcaResult aca_partial_pivoting(const KGen &g, double epsilon) {  // aca.rs:23-136
  const int nr = g.P, ncol = g.P;
  const int max_it = std::min(nr, ncol);
  const double tol = epsilon * epsilon;
  std::vector<uint8_t> unused_rows(nr, 1), unused_cols(ncol, 1);
  std::vector<std::vector<double>> us, vs;
  std::vector<double> row(ncol), col(nr);
  double residual_norm = 0.0, sum_k = 0.0;
  int i = 0, k = 0;
  for (int it = 0; it < max_it; ++it) {
    for (int j = 0; j < ncol; ++j) row[j] = g.at(i, j);
    unused_rows[i] = 0;
    for (int l = 0; l < k; ++l) {
      const double ui = us[l][i];
      for (int j = 0; j < ncol; ++j) row[j] -= ui * vs[l][j];
    }
    const int j = argmax_masked(row, unused_cols);
    const double pivot = 1.0 / row[j];
    for (int jj = 0; jj < ncol; ++jj) row[jj] *= pivot;
    for (int ii = 0; ii < nr; ++ii) col[ii] = g.at(ii, j);
    unused_cols[j] = 0;
    for (int l = 0; l < k; ++l) {
      const double vj = vs[l][j];
      for (int ii = 0; ii < nr; ++ii) col[ii] -= vj * us[l][ii];
    }
    i = argmax_masked(col, unused_rows);
    if (k > 0) {
      sum_k = 0.0;
      for (int l = 0; l < k; ++l) {
        double a = 0, b = 0;
        for (int ii = 0; ii < nr; ++ii) a += us[l][ii] * col[ii];
        for (int jj = 0; jj < ncol; ++jj) b += vs[l][jj] * row[jj];
        sum_k += a * b;
      }
    }
    double nu = 0, nv = 0;
    for (int ii = 0; ii < nr; ++ii) nu += col[ii] * col[ii];
    for (int jj = 0; jj < ncol; ++jj) nv += row[jj] * row[jj];
    const double norm_u_v_2 = nu * nv;
    residual_norm += norm_u_v_2 + 2.0 * sum_k;
    us.push_back(col);
    vs.push_back(row);
    k += 1;
    if (norm_u_v_2 <= tol * residual_norm) break;
  }
  AcaResult r;
  r.u = Mat(nr, k);
  r.v = Mat(ncol, k);
  for (int l = 0; l < k; ++l) {
    std::copy(us[l].begin(), us[l].end(), r.u.a.begin() + (size_t)l * nr);
    std::copy(vs[l].begin(), vs[l].end(), r.v.a.begin() + (size_t)l * ncol);
  }
  return r;
}

Mat matmul(const Mat &A, const Mat &B) {
  Mat C(A.rows, B.cols);
  for (int j = 0; j < B.cols; ++j)
    for (int k = 0; k < A.cols; ++k) {
      const double b = B(k, j);
      if (b == 0.0) continue;
      for (int i = 0; i < A.rows; ++i) C(i, j) += A(i, k) * b;
    }
  return C;
}

Mat transpose(const Mat &A) {
  Mat T(A.cols, A.rows);
  for (int j = 0; j < A.cols; ++j)
    for (int i = 0; i < A.rows; ++i) T(j, i) = A(i, j);
  return T;
}

void recompress_aca(const AcaResult &aca, double eps, M2LOperator &op) {  // aca.rs:173-200
  Mat qu, ru, qv, rv;
  thin_qr(aca.u, qu, ru);
  thin_qr(aca.v, qv, rv);
  Mat core = matmul(ru, transpose(rv));
  Mat ur, vr;
  std::vector<double> sr;
  jacobi_svd(core, ur, sr, vr);
  const int rank = singular_values_cutoff(sr, eps);
  Mat us(core.rows, rank);
  for (int j = 0; j < rank; ++j)
    for (int i = 0; i < core.rows; ++i) us(i, j) = ur(i, j) * sr[j];
  op.rank = rank;
  op.U = matmul(qu, us);                      // P x rank
  Mat vrt(rank, core.cols);                   // V_r^T rows
  for (int j = 0; j < rank; ++j)
    for (int i = 0; i < core.cols; ++i) vrt(j, i) = vr(i, j);
  op.Vt = matmul(vrt, transpose(qv));         // rank x P
}

std::vector<int> argsort_stable(const std::vector<int> &d) {
  std::vector<int> idx(d.size());
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return d[a] < d[b]; });
  return idx;
}

}  // namespace

void Operators::build(int order, int dim_, double radius, int depth, const KParams &kp, int compression,
                      double eps) {
  p = order;
  dim = dim_;
  P = 1;
  for (int d = 0; d < dim; ++d) P *= p;
  nodes.resize(p);
  for (int k = 0; k < p; ++k) {  // chebyshev.rs:32-40: i = p-1 ... 0
    const int i = p - 1 - k;
    nodes[k] = std::cos(M_PI * ((double)i + 0.5) / (double)p);
  }
  tnodes.assign((size_t)p * p, 0.0);
  for (int m = 0; m < p; ++m) cheb_T(p, nodes[m], &tnodes[(size_t)m * p]);
  // child transfer S (chebyshev.rs:146-180)
  child_s.assign((size_t)2 * p * p, 0.0);
  std::vector<double> T(p);
  for (int h = 0; h < 2; ++h)
    for (int i = 0; i < p; ++i) {
      const double x = (h == 0 ? nodes[i] - 1.0 : nodes[i] + 1.0) * 0.5;
      cheb_T(p, x, T.data());
      for (int m = 0; m < p; ++m) {
        double s = 0;
        for (int k = 0; k < p; ++k) s += T[k] * tnodes[(size_t)m * p + k];
        child_s[((size_t)h * p + i) * p + m] = (s * 2.0 - 1.0) / (double)p;  // chebyshev.rs:114-127
      }
    }

  // ---- symmetry tables (chebyshev.rs:267-585)
  n_vec = 1;
  for (int d = 0; d < dim; ++d) n_vec *= 7;
  auto vec_of = [&](int idx, int base, int lo, int *v) {  // cartesian_product: column 0 slowest
    for (int d = dim - 1; d >= 0; --d) {
      v[d] = lo + idx % base;
      idx /= base;
    }
  };
  ref_vecs.clear();
  {
    int nb = 1;
    for (int d = 0; d < dim; ++d) nb *= 4;
    for (int i = 0; i < nb; ++i) {
      int v[3] = {0, 0, 0};
      vec_of(i, 4, 0, v);
      bool ok = v[0] >= 2;
      for (int d = 1; d < dim && ok; ++d)
        if (v[d] > v[d - 1]) ok = false;
      if (ok)
        for (int d = 0; d < dim; ++d) ref_vecs.push_back(v[d]);
    }
  }
  n_ref = (int)ref_vecs.size() / dim;
  std::vector<std::vector<int>> axis_orders;
  {
    std::vector<int> o(dim);
    std::iota(o.begin(), o.end(), 0);
    do axis_orders.push_back(o);
    while (std::next_permutation(o.begin(), o.end()));
  }
  const int n_sign = 1 << dim, n_order = (int)axis_orders.size();
  n_perm = n_sign * n_order;
  std::vector<int> pw(dim);
  for (int d = 0; d < dim; ++d) {
    pw[d] = 1;
    for (int e = d + 1; e < dim; ++e) pw[d] *= p;
  }
  auto multi = [&](int j, int *alpha) {  // 1-based multi-index, column 0 slowest
    for (int d = dim - 1; d >= 0; --d) {
      alpha[d] = 1 + j % p;
      j /= p;
    }
  };
  auto to_k = [&](const int *alpha) {
    int k = 0;
    for (int d = 0; d < dim; ++d) k += (alpha[d] - 1) * pw[d];
    return k;
  };
  std::vector<std::vector<int>> diag(n_order, std::vector<int>(P)), axial(n_sign, std::vector<int>(P));
  for (int o = 0; o < n_order; ++o)
    for (int j = 0; j < P; ++j) {
      int a[3], ap[3];
      multi(j, a);
      for (int d = 0; d < dim; ++d) ap[d] = a[axis_orders[o][d]];
      diag[o][to_k(ap)] = j;
    }
  for (int s = 0; s < n_sign; ++s) {
    int sg[3];
    vec_of(s, 2, 0, sg);  // 0 -> -1, 1 -> +1 (cartesian_product([-1, 1]))
    for (int j = 0; j < P; ++j) {
      int a[3], ap[3];
      multi(j, a);
      for (int d = 0; d < dim; ++d) ap[d] = sg[d] == 0 ? p - (a[d] - 1) : a[d];
      axial[s][to_k(ap)] = j;
    }
  }
  perm.assign((size_t)n_perm * P, 0);
  inv_perm.assign((size_t)n_perm * P, 0);
  for (int a = 0; a < n_sign; ++a)
    for (int b = 0; b < n_order; ++b) {
      const int id = a * n_order + b;
      std::vector<int> comb(P);
      for (int i = 0; i < P; ++i) comb[i] = axial[a][diag[b][i]];
      std::vector<int> inv = argsort_stable(comb);
      for (int i = 0; i < P; ++i) {
        perm[(size_t)id * P + i] = comb[i];
        inv_perm[(size_t)id * P + i] = inv[i];
      }
    }
  perm_lookup.assign(n_vec, 0);
  ref_lookup.assign(n_vec, 0);
  for (int t = 0; t < n_vec; ++t) {
    int v[3] = {0, 0, 0};
    vec_of(t, 7, -3, v);
    int a = 0;
    for (int d = 0; d < dim; ++d) a = a * 2 + (v[d] < 0 ? 0 : 1);
    std::vector<int> negabs(dim);
    for (int d = 0; d < dim; ++d) negabs[d] = -std::abs(v[d]);
    std::vector<int> so = argsort_stable(negabs);
    int b = 0;
    for (int o = 0; o < n_order; ++o)
      if (axis_orders[o] == so) {
        b = o;
        break;
      }
    perm_lookup[t] = a * n_order + b;
    std::vector<int> sv(dim);
    for (int d = 0; d < dim; ++d) sv[d] = std::abs(v[d]);
    std::sort(sv.begin(), sv.end());
    for (int r = 0; r < n_ref; ++r) {
      std::vector<int> rv(ref_vecs.begin() + (size_t)r * dim, ref_vecs.begin() + (size_t)(r + 1) * dim);
      std::sort(rv.begin(), rv.end());
      if (rv == sv) {
        ref_lookup[t] = r;
        break;
      }
    }
  }

  // ---- M2L operators per level 2..depth and reference vector (chebyshev.rs:697-791)
  std::vector<double> nodes_nd((size_t)P * dim);
  for (int j = 0; j < P; ++j) {
    int a[3];
    multi(j, a);
    for (int d = 0; d < dim; ++d) nodes_nd[(size_t)j * dim + d] = nodes[a[d] - 1];
  }
  const int n_levels = depth >= 2 ? depth - 1 : 0;
  m2l.assign(n_levels, std::vector<M2LOperator>(n_ref));
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
  for (int li = 0; li < n_levels; ++li)
    for (int r = 0; r < n_ref; ++r) {
      const int lvl = li + 2;
      const double cell_length = radius / (double)(1ull << (lvl - 1));  // chebyshev.rs:702
      std::vector<double> tgt((size_t)P * dim), src((size_t)P * dim);
      for (int j = 0; j < P; ++j)
        for (int d = 0; d < dim; ++d) {
          const double x = nodes_nd[(size_t)j * dim + d];
          tgt[(size_t)j * dim + d] = x * (0.5 * cell_length);                                        // :588-600
          src[(size_t)j * dim + d] = ((double)ref_vecs[(size_t)r * dim + d] + (x * 0.5)) * cell_length;  // :604-627
        }
      KGen g{&src, &tgt, P, dim, &kp};
      M2LOperator &op = m2l[li][r];
      if (compression == FB_COMPRESSION_ACA) {
        AcaResult aca = aca_partial_pivoting(g, eps);
        recompress_aca(aca, eps, op);
      } else {
        Mat K(P, P);
        for (int j = 0; j < P; ++j)
          for (int i = 0; i < P; ++i) K(i, j) = g.at(i, j);
        if (compression == FB_COMPRESSION_SVD) {  // chebyshev.rs:760-779
          Mat ur, vr;
          std::vector<double> sr;
          jacobi_svd(K, ur, sr, vr);
          const int rank = singular_values_cutoff(sr, eps);
          op.rank = rank;
          op.U = Mat(P, rank);
          op.Vt = Mat(rank, P);
          for (int j = 0; j < rank; ++j)
            for (int i = 0; i < P; ++i) {
              op.U(i, j) = ur(i, j);
              op.Vt(j, i) = sr[j] * vr(i, j);
            }
        } else {
          op.rank = P;
          op.U = K;
        }
      }
    }
}

}  // namespace fb

// ---------------------------------------------------------------------------------------- operator cache
#include <cstring>
#include <list>
#include <mutex>

namespace fb {

size_t Operators::bytes() const {
  size_t b = sizeof(double) * (nodes.size() + tnodes.size() + child_s.size()) +
             sizeof(int32_t) * (perm.size() + inv_perm.size() + perm_lookup.size() + ref_lookup.size() + ref_vecs.size());
  for (const auto &lv : m2l)
    for (const auto &op : lv) b += sizeof(double) * (op.U.a.size() + op.Vt.a.size());
  return b;
}

namespace {
struct OpsKey {
  int order, dim, depth, compression;
  double radius, eps;
  KParams kp;
  bool operator==(const OpsKey &o) const {
    return order == o.order && dim == o.dim && depth == o.depth && compression == o.compression &&
           std::memcmp(&radius, &o.radius, sizeof(double)) == 0 && std::memcmp(&eps, &o.eps, sizeof(double)) == 0 &&
           kp.fam == o.kp.fam && kp.pw == o.kp.pw && std::memcmp(&kp.s2, &o.kp.s2, sizeof(double)) == 0 &&
           std::memcmp(&kp.ip2, &o.kp.ip2, sizeof(double)) == 0 &&
           std::memcmp(&kp.near_slope, &o.kp.near_slope, sizeof(double)) == 0 &&
           std::memcmp(&kp.far_coef, &o.kp.far_coef, sizeof(double)) == 0 &&
           std::memcmp(&kp.total_sill, &o.kp.total_sill, sizeof(double)) == 0;
  }
};
std::mutex g_ops_mutex;
std::list<std::pair<OpsKey, Operators>> g_ops_cache;  // most recently used first
constexpr size_t kOpsCacheEntries = 4;
constexpr size_t kOpsCacheBytes = (size_t)768 << 20;
}  // namespace

bool Operators::build_cached(int order, int dim_, double radius, int depth, const KParams &kp, int compression,
                             double eps) {
  // kp.fast selects the square-root variant of the device loops only: the host operators do not depend on it
  const OpsKey key{order, dim_, depth, compression, radius, eps, kp};
  {
    std::lock_guard<std::mutex> lock(g_ops_mutex);
    for (auto it = g_ops_cache.begin(); it != g_ops_cache.end(); ++it)
      if (it->first == key) {
        *this = it->second;
        g_ops_cache.splice(g_ops_cache.begin(), g_ops_cache, it);
        return true;
      }
  }
  build(order, dim_, radius, depth, kp, compression, eps);
  if (bytes() <= kOpsCacheBytes / 2) {
    std::lock_guard<std::mutex> lock(g_ops_mutex);
    g_ops_cache.emplace_front(key, *this);
    size_t total = 0;
    for (const auto &e : g_ops_cache) total += e.second.bytes();
    while (g_ops_cache.size() > kOpsCacheEntries || (total > kOpsCacheBytes && g_ops_cache.size() > 1)) {
      total -= g_ops_cache.back().second.bytes();
      g_ops_cache.pop_back();
    }
  }
  return false;
}

}  // namespace fb

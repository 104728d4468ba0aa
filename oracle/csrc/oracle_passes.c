/*
 * oracle_passes.c — C/OpenMP restatement of the reference's hot loops, used by the numpy oracle
 * for larger cases and by bench.py's cpu_baseline / --impl reference legs.
 *
 * TEST INFRASTRUCTURE ONLY: never linked into or loaded by the product package.
 *
 *   orc_direct        particle_to_particle / multipole_to_particle / particle_to_local inner loops
 *                     (ferreus_bbfmm/src/bbfmm.rs:1162-1355, 1001-1048): out[t,r] += k(x_t, y_s) w[s,r]
 *   orc_kernel_value  the kernel zoo (ferreus_rbf_utils/src/rbf_kernels.rs:25-317,
 *                     non_rbf_kernels.rs:20-163)
 *   orc_leaf_pass     the rayon loop over target leaves (bbfmm.rs:1113-1159) for P2P + M2P
 *   orc_m2l           multipole_to_local (bbfmm.rs:864-986) over all target cells of all levels
 *
 * Like the reference, the kernel is re-evaluated for every right-hand side (bbfmm.rs:1184).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EPS 2.220446049250313e-16

typedef struct {
  int kernel_type; /* registry order, utils.rs:558-571 */
  int pw;
  double s2, ip2, near_slope, far_coef, total_sill;
} orc_kernel;

static inline double kval(const orc_kernel *k, double r2) {
  switch (k->kernel_type) {
    case 0: return -sqrt(r2);
    case 1: { double r = sqrt(r2); return fabs(r) < EPS ? 0.0 : (r * r) * log(r); }
    case 2: { double r = sqrt(r2); return r * r * r; }
    case 3: case 4: case 5: case 6: {
      double sr2 = k->s2 * r2;
      if (sr2 <= k->ip2) return k->total_sill - k->near_slope * sqrt(r2);
      double t = 1.0 + sr2, t2 = t * t;
      double tp = k->pw == 1 ? t : (k->pw == 2 ? t2 : (k->pw == 3 ? t2 * t : t2 * t2));
      return k->far_coef / (tp * sqrt(t));
    }
    case 7: { double r = sqrt(r2); return fabs(r) < EPS ? 0.0 : 1.0 / r; }
    case 8: { double r = sqrt(r2); return fabs(r) < EPS ? 0.0 : 1.0 / (r * r); }
    default: { double r = sqrt(r2); double rr = r * r; return fabs(r) < EPS ? 0.0 : 1.0 / (rr * rr); }
  }
}

double orc_kernel_value(const orc_kernel *k, double r2) { return kval(k, r2); }

void orc_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* targets nt x dim (row-major), sources ns x dim, weights ns x nrhs (row stride w_rs), out nt x nrhs (row stride o_rs) */
static void direct_block(const orc_kernel *k, int dim, const double *tg, int nt, const double *src, int ns,
                         const double *w, int w_rs, int nrhs, double *out, int o_rs) {
  for (int r = 0; r < nrhs; ++r)          /* bbfmm.rs:1184: per right-hand side */
    for (int t = 0; t < nt; ++t) {
      const double *x = tg + (size_t)t * dim;
      double acc = 0.0;
      for (int s = 0; s < ns; ++s) {
        const double *y = src + (size_t)s * dim;
        double r2 = 0.0;
        for (int d = 0; d < dim; ++d) { double df = x[d] - y[d]; r2 += df * df; }
        acc += kval(k, r2) * w[(size_t)s * w_rs + r];
      }
      out[(size_t)t * o_rs + r] += acc;
    }
}

void orc_direct(const orc_kernel *k, int dim, const double *tg, int nt, const double *src, int ns, const double *w,
                int w_rs, int nrhs, double *out, int o_rs) {
  direct_block(k, dim, tg, nt, src, ns, w, w_rs, nrhs, out, o_rs);
}

/*
 * Leaf pass over a list of target leaves (P2P + M2P).  All index arrays are CSR:
 *   leaf l: targets t_idx[t_ptr[l]..t_ptr[l+1]) (rows of `targets`),
 *           U sources  u_idx[u_ptr[l]..) (rows of `sources`, already concatenated over the U cells),
 *           W cells    w_cell[w_ptr[l]..) with centres wc (n_wcells x dim), half sides wh, multipoles
 *           mult[(cell * nrhs + r) * P + node] and reference nodes nodes_nd (P x dim in [-1,1]).
 * out: m x nrhs row-major, += .
 */
void orc_leaf_pass(const orc_kernel *k, int dim, int nrhs, int n_leaves, const int64_t *t_ptr, const int64_t *t_idx,
                   const double *targets, const int64_t *u_ptr, const int64_t *u_idx, const double *sources,
                   const double *weights, const int64_t *w_ptr, const int64_t *w_cell, const double *cell_center,
                   const double *cell_half, const double *mult, int P, const double *nodes_nd, double *out) {
#pragma omp parallel
  {
    double *tbuf = NULL, *sbuf = NULL, *wbuf = NULL, *obuf = NULL;
    size_t tcap = 0, scap = 0;
#pragma omp for schedule(dynamic, 1)
    for (int l = 0; l < n_leaves; ++l) {
      const int nt = (int)(t_ptr[l + 1] - t_ptr[l]);
      if (nt == 0) continue;
      if ((size_t)nt > tcap) {
        tcap = (size_t)nt * 2;
        tbuf = (double *)realloc(tbuf, tcap * dim * sizeof(double));
        obuf = (double *)realloc(obuf, tcap * nrhs * sizeof(double));
      }
      for (int t = 0; t < nt; ++t)
        memcpy(tbuf + (size_t)t * dim, targets + (size_t)t_idx[t_ptr[l] + t] * dim, dim * sizeof(double));
      memset(obuf, 0, (size_t)nt * nrhs * sizeof(double));
      const int ns = (int)(u_ptr[l + 1] - u_ptr[l]);
      size_t need = (size_t)(ns > P ? ns : P);
      if (need > scap) {
        scap = need * 2;
        sbuf = (double *)realloc(sbuf, scap * dim * sizeof(double));
        wbuf = (double *)realloc(wbuf, scap * nrhs * sizeof(double));
      }
      for (int s = 0; s < ns; ++s) {
        const int64_t g = u_idx[u_ptr[l] + s];
        memcpy(sbuf + (size_t)s * dim, sources + (size_t)g * dim, dim * sizeof(double));
        memcpy(wbuf + (size_t)s * nrhs, weights + (size_t)g * nrhs, nrhs * sizeof(double));
      }
      direct_block(k, dim, tbuf, nt, sbuf, ns, wbuf, nrhs, nrhs, obuf, nrhs);
      for (int64_t e = w_ptr[l]; e < w_ptr[l + 1]; ++e) {
        const int64_t c = w_cell[e];
        for (int j = 0; j < P; ++j) {
          for (int d = 0; d < dim; ++d)
            sbuf[(size_t)j * dim + d] = cell_center[c * dim + d] + cell_half[c] * nodes_nd[(size_t)j * dim + d];
          for (int r = 0; r < nrhs; ++r) wbuf[(size_t)j * nrhs + r] = mult[((size_t)c * nrhs + r) * P + j];
        }
        direct_block(k, dim, tbuf, nt, sbuf, P, wbuf, nrhs, nrhs, obuf, nrhs);
      }
      for (int t = 0; t < nt; ++t)
        for (int r = 0; r < nrhs; ++r) out[(size_t)t_idx[t_ptr[l] + t] * nrhs + r] += obuf[(size_t)t * nrhs + r];
    }
    free(tbuf); free(sbuf); free(wbuf); free(obuf);
  }
}

/*
 * P2L for a list of cells: loc[(cell*nrhs + r)*P + node] += sum_s k(node, y_s) w[s, r]  (bbfmm.rs:1001-1048)
 */
void orc_p2l(const orc_kernel *k, int dim, int nrhs, int n_cells, const int64_t *cells, const int64_t *x_ptr,
             const int64_t *x_idx, const double *sources, const double *weights, const double *cell_center,
             const double *cell_half, int P, const double *nodes_nd, double *loc) {
#pragma omp parallel
  {
    double *nbuf = (double *)malloc((size_t)P * dim * sizeof(double));
    double *obuf = (double *)malloc((size_t)P * nrhs * sizeof(double));
    double *sbuf = NULL, *wbuf = NULL;
    size_t scap = 0;
#pragma omp for schedule(dynamic, 1)
    for (int i = 0; i < n_cells; ++i) {
      const int64_t c = cells[i];
      const int ns = (int)(x_ptr[i + 1] - x_ptr[i]);
      if (ns == 0) continue;
      if ((size_t)ns > scap) {
        scap = (size_t)ns * 2;
        sbuf = (double *)realloc(sbuf, scap * dim * sizeof(double));
        wbuf = (double *)realloc(wbuf, scap * nrhs * sizeof(double));
      }
      for (int j = 0; j < P; ++j)
        for (int d = 0; d < dim; ++d)
          nbuf[(size_t)j * dim + d] = cell_center[c * dim + d] + cell_half[c] * nodes_nd[(size_t)j * dim + d];
      for (int s = 0; s < ns; ++s) {
        const int64_t g = x_idx[x_ptr[i] + s];
        memcpy(sbuf + (size_t)s * dim, sources + (size_t)g * dim, dim * sizeof(double));
        memcpy(wbuf + (size_t)s * nrhs, weights + (size_t)g * nrhs, nrhs * sizeof(double));
      }
      memset(obuf, 0, (size_t)P * nrhs * sizeof(double));
      direct_block(k, dim, nbuf, P, sbuf, ns, wbuf, nrhs, nrhs, obuf, nrhs);
      for (int r = 0; r < nrhs; ++r)
        for (int j = 0; j < P; ++j) loc[((size_t)c * nrhs + r) * P + j] += obuf[(size_t)j * nrhs + r];
    }
    free(nbuf); free(obuf); free(sbuf); free(wbuf);
  }
}

/*
 * M2L over entries grouped by target cell (CSR e_ptr over n_tgt target cells):
 *   x = M_src[perm], y = U (Vt x) (or K x), L_tgt[i] += y[inv_perm[i]]     (bbfmm.rs:864-986)
 * Operators: op_u[op_id] (P x rank, column-major), op_vt[op_id] (rank x P, column-major) or NULL.
 */
void orc_m2l(int P, int nrhs, int n_tgt, const int64_t *tgt_cell, const int64_t *e_ptr, const int64_t *e_src,
             const int32_t *e_op, const int32_t *e_perm, const int32_t *perm_tab, const int32_t *inv_tab,
             const double *const *op_u, const double *const *op_vt, const int32_t *op_rank, const double *mult,
             double *loc) {
#pragma omp parallel
  {
    double *x = (double *)malloc((size_t)P * sizeof(double));
    double *y = (double *)malloc((size_t)P * sizeof(double));
    double *z = (double *)malloc((size_t)P * sizeof(double));
#pragma omp for schedule(dynamic, 4)
    for (int i = 0; i < n_tgt; ++i) {
      const int64_t c = tgt_cell[i];
      for (int64_t e = e_ptr[i]; e < e_ptr[i + 1]; ++e) {
        const int op = e_op[e];
        const int rk = op_rank[op];
        const int32_t *pm = perm_tab + (size_t)e_perm[e] * P;
        const int32_t *iv = inv_tab + (size_t)e_perm[e] * P;
        for (int r = 0; r < nrhs; ++r) {
          const double *m = mult + ((size_t)e_src[e] * nrhs + r) * P;
          for (int j = 0; j < P; ++j) x[j] = m[pm[j]];
          const double *U = op_u[op], *Vt = op_vt[op];
          if (Vt) {
            for (int q = 0; q < rk; ++q) y[q] = 0.0;
            for (int j = 0; j < P; ++j) {
              const double xj = x[j];
              const double *col = Vt + (size_t)j * rk;
              for (int q = 0; q < rk; ++q) y[q] += col[q] * xj;
            }
            for (int j = 0; j < P; ++j) z[j] = 0.0;
            for (int q = 0; q < rk; ++q) {
              const double yq = y[q];
              const double *col = U + (size_t)q * P;
              for (int j = 0; j < P; ++j) z[j] += col[j] * yq;
            }
          } else {
            for (int j = 0; j < P; ++j) z[j] = 0.0;
            for (int q = 0; q < P; ++q) {
              const double xq = x[q];
              const double *col = U + (size_t)q * P;
              for (int j = 0; j < P; ++j) z[j] += col[j] * xq;
            }
          }
          double *L = loc + ((size_t)c * nrhs + r) * P;
          for (int j = 0; j < P; ++j) L[j] += z[iv[j]];
        }
      }
    }
    free(x); free(y); free(z);
  }
}

/* ---- Chebyshev transfers (added round 2: the upward pass, L2L and L2P were Python loops over cells) ---------------
 * S_n(x)[m] = (2 sum_k T_k(x) T_k(x_m) - 1) / p   (chebyshev.rs:114-127), T by the three-term recurrence (:47-110);
 * tensor weights with axis 0 slowest (chebyshev.rs:894-906).  tn[m * p + k] = T_k(node_m).                         */
static void cheb_sn(int p, double x, const double *tn, double *s) {
  double T[32];
  T[0] = 1.0;
  if (p > 1) T[1] = x;
  for (int k = 2; k < p; ++k) T[k] = 2.0 * x * T[k - 1] - T[k - 2];
  for (int m = 0; m < p; ++m) {
    double acc = 0.0;
    for (int k = 0; k < p; ++k) acc += T[k] * tn[m * p + k];
    s[m] = (acc * 2.0 - 1.0) / (double)p;
  }
}

static void tensor_weights(int p, int dim, const double *pt, const double *center, double half, const double *tn,
                           double *S /* p^dim */) {
  double s[3][32];
  for (int d = 0; d < dim; ++d) cheb_sn(p, (pt[d] - center[d]) / half, tn, s[d]);
  if (dim == 1) {
    for (int i = 0; i < p; ++i) S[i] = s[0][i];
  } else if (dim == 2) {
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < p; ++j) S[i * p + j] = s[0][i] * s[1][j];
  } else {
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < p; ++j) {
        const double sij = s[0][i] * s[1][j];
        for (int k = 0; k < p; ++k) S[(i * p + j) * p + k] = sij * s[2][k];
      }
  }
}

/* P2M (bbfmm.rs:691-739): M[cell][rhs][:] += sum_{points of the leaf} S(point) w[point][rhs] */
void orc_p2m(int p, int dim, int P, int nrhs, int n_leaves, const int64_t *leaf_cell, const int64_t *t_ptr,
             const int64_t *t_idx, const double *src, const double *w, const double *center, const double *half,
             const double *tn, double *M) {
#pragma omp parallel
  {
    double *S = (double *)malloc(sizeof(double) * (size_t)P);
#pragma omp for schedule(dynamic, 4)
    for (int l = 0; l < n_leaves; ++l) {
      const int64_t c = leaf_cell[l];
      double *Mc = M + (size_t)c * nrhs * P;
      for (int64_t e = t_ptr[l]; e < t_ptr[l + 1]; ++e) {
        const int64_t i = t_idx[e];
        tensor_weights(p, dim, src + (size_t)i * dim, center + (size_t)c * dim, half[c], tn, S);
        for (int r = 0; r < nrhs; ++r) {
          const double wr = w[(size_t)i * nrhs + r];
          double *Mr = Mc + (size_t)r * P;
          for (int k = 0; k < P; ++k) Mr[k] += S[k] * wr;
        }
      }
    }
    free(S);
  }
}

/* L2P (bbfmm.rs:1358-1440, values): out[point][rhs] += S(point) . L[cell][rhs][:] */
void orc_l2p(int p, int dim, int P, int nrhs, int n_leaves, const int64_t *leaf_cell, const int64_t *t_ptr,
             const int64_t *t_idx, const double *src, const double *center, const double *half, const double *tn,
             const double *L, double *out) {
#pragma omp parallel
  {
    double *S = (double *)malloc(sizeof(double) * (size_t)P);
#pragma omp for schedule(dynamic, 4)
    for (int l = 0; l < n_leaves; ++l) {
      const int64_t c = leaf_cell[l];
      const double *Lc = L + (size_t)c * nrhs * P;
      for (int64_t e = t_ptr[l]; e < t_ptr[l + 1]; ++e) {
        const int64_t i = t_idx[e];
        tensor_weights(p, dim, src + (size_t)i * dim, center + (size_t)c * dim, half[c], tn, S);
        for (int r = 0; r < nrhs; ++r) {
          const double *Lr = Lc + (size_t)r * P;
          double acc = 0.0;
          for (int k = 0; k < P; ++k) acc += S[k] * Lr[k];
          out[(size_t)i * nrhs + r] += acc;
        }
      }
    }
    free(S);
  }
}

/* one level of M2M (bbfmm.rs:742-772): M[parent] += M2M[slot(child)] M[child]; parents of the level are independent.
 * one level of L2L (bbfmm.rs:1051-1086): L[child] += M2M[slot(child)]^T L[parent]; children are independent.
 * m2m: [2^dim][P][P] row-major; up != 0 selects M2M.                                                                */
void orc_transfer_level(int P, int nrhs, int up, int n_parents, const int64_t *parents, const int64_t *child_ptr,
                        const int64_t *child_idx, const int64_t *child_slot, const double *m2m, double *X) {
#pragma omp parallel for schedule(dynamic, 2)
  for (int q = 0; q < n_parents; ++q) {
    const int64_t par = parents[q];
    for (int64_t e = child_ptr[q]; e < child_ptr[q + 1]; ++e) {
      const int64_t ch = child_idx[e];
      const double *A = m2m + (size_t)child_slot[ch] * P * P;
      for (int r = 0; r < nrhs; ++r) {
        double *xp = X + ((size_t)par * nrhs + r) * P, *xc = X + ((size_t)ch * nrhs + r) * P;
        if (up) {
          for (int i = 0; i < P; ++i) {
            double acc = 0.0;
            const double *row = A + (size_t)i * P;
            for (int j = 0; j < P; ++j) acc += row[j] * xc[j];
            xp[i] += acc;
          }
        } else {
          for (int i = 0; i < P; ++i) {
            const double v = xp[i];
            const double *row = A + (size_t)i * P;
            for (int j = 0; j < P; ++j) xc[j] += row[j] * v;
          }
        }
      }
    }
  }
}

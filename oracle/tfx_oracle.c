/*
 * tfx_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the Tomofast-x inversion hot path. See tfx_oracle.h
 * for the pinning statement. Build: `make -C oracle` (gcc -O3 -ffp-contract=off,
 * mirroring the reference's `gfortran -O3` on baseline x86-64: no FMA contraction).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
 * reference) may load the resulting library.
 */
#include "tfx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ========================================================================== */
/* t_sparse_matrix -- src/inversion/sparse_matrix.f90                          */
/* ========================================================================== */

/* sparse_matrix_initialize + allocate_arrays, sparse_matrix.f90:105-130,498-526 */
orc_csr *orc_csr_new(int32_t nl, int32_t ncolumns, int64_t nnz, int32_t nl_empty) {
  if (nnz < 0 || nl < 0) return NULL;
  orc_csr *m = (orc_csr *)calloc(1, sizeof(orc_csr));
  m->nl = nl;
  m->ncolumns = ncolumns;
  m->nnz = nnz;
  m->nl_nonempty_allocated = nl - nl_empty;
  m->sa = (float *)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(float));
  m->ija = (int32_t *)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(int32_t));
  m->ijl = (int64_t *)calloc((size_t)m->nl_nonempty_allocated + 1, sizeof(int64_t));
  m->rowptr = (int32_t *)calloc((size_t)(m->nl_nonempty_allocated > 0 ? m->nl_nonempty_allocated : 1),
                                sizeof(int32_t));
  return m;
}

void orc_csr_free(orc_csr *m) {
  if (!m) return;
  free(m->sa); free(m->ija); free(m->ijl); free(m->rowptr); free(m);
}

/* sparse_matrix_reset, sparse_matrix.f90:135-151 */
void orc_csr_reset(orc_csr *m) {
  m->nl_current = 0; m->nl_current_all = 0; m->nel = 0; m->nel_last = 0; m->nl_nonempty = 0;
  m->finalized = 0;
  memset(m->sa, 0, sizeof(float) * (size_t)(m->nnz > 0 ? m->nnz : 1));
  memset(m->ija, 0, sizeof(int32_t) * (size_t)(m->nnz > 0 ? m->nnz : 1));
  memset(m->ijl, 0, sizeof(int64_t) * ((size_t)m->nl_nonempty_allocated + 1));
  memset(m->rowptr, 0, sizeof(int32_t) * (size_t)(m->nl_nonempty_allocated > 0 ? m->nl_nonempty_allocated : 1));
}

/* sparse_matrix_add, sparse_matrix.f90:213-229: zero values are never stored,
 * the value is rounded to real(4). */
int orc_csr_add(orc_csr *m, double value, int32_t column) {
  if (value == 0.0) return 0;
  if (m->nel >= m->nnz) return -1;
  m->sa[m->nel] = (float)value;
  m->ija[m->nel] = column;
  m->nel += 1;
  return 0;
}

/* sparse_matrix_add_row, sparse_matrix.f90:234-249 */
int orc_csr_add_row(orc_csr *m, int32_t nel_add, const float *values, const int32_t *columns) {
  if (m->nel + nel_add > m->nnz) return -1;
  memcpy(m->sa + m->nel, values, sizeof(float) * (size_t)nel_add);
  memcpy(m->ija + m->nel, columns, sizeof(int32_t) * (size_t)nel_add);
  m->nel += nel_add;
  return 0;
}

/* sparse_matrix_new_row, sparse_matrix.f90:254-276: only non-empty rows are stored */
int orc_csr_new_row(orc_csr *m) {
  if (m->nl_current >= m->nl_nonempty_allocated) return -1;
  m->nl_current_all += 1;
  if (m->nel > m->nel_last) {
    m->nl_current += 1;
    m->ijl[m->nl_current - 1] = m->nel_last + 1;
    m->rowptr[m->nl_current - 1] = m->nl_current_all;
    m->nel_last = m->nel;
  }
  return 0;
}

/* sparse_matrix_add_empty_rows, sparse_matrix.f90:281-293 */
void orc_csr_add_empty_rows(orc_csr *m, int32_t nrows) { m->nl_current_all += nrows; }

/* sparse_matrix_finalize + validate, sparse_matrix.f90:157-208 */
int orc_csr_finalize(orc_csr *m) {
  if (m->nl_current_all != m->nl) return -1;
  if (m->nel_last != m->nel) return -2;
  m->ijl[m->nl_current] = m->nel + 1;
  m->nl_nonempty = m->nl_current;
  for (int32_t i = 0; i < m->nl_nonempty; ++i) {
    for (int64_t k = m->ijl[i]; k <= m->ijl[i + 1] - 1; ++k) {
      if (k < 1 || k > m->nnz) return -3;
      int32_t j = m->ija[k - 1];
      if (j < 1 || j > m->ncolumns) return -4;
    }
  }
  m->finalized = 1;
  return 0;
}

/* sparse_matrix_add_mult_vector, sparse_matrix.f90:313-329 (sequential sum per row) */
void orc_csr_add_mult_vector(const orc_csr *m, const double *x, double *b) {
  for (int32_t i = 0; i < m->nl_nonempty; ++i) {
    int32_t i_all = m->rowptr[i] - 1;
    for (int64_t k = m->ijl[i] - 1; k < m->ijl[i + 1] - 1; ++k)
      b[i_all] = b[i_all] + (double)m->sa[k] * x[m->ija[k] - 1];
  }
}

/* sparse_matrix_mult_vector, sparse_matrix.f90:298-307 */
void orc_csr_mult_vector(const orc_csr *m, const double *x, double *b) {
  for (int32_t i = 0; i < m->nl; ++i) b[i] = 0.0;
  orc_csr_add_mult_vector(m, x, b);
}

/* sparse_matrix_part_mult_vector, sparse_matrix.f90:335-367 */
int orc_csr_part_mult_vector(const orc_csr *m, int32_t nelements, const double *x, int32_t ndata,
                             double *b, int32_t line_start, int32_t param_shift) {
  (void)nelements;
  int32_t line_end = line_start + ndata - 1;
  if (line_start < 1 || line_start > m->nl_current_all || line_end < 1 || line_end > m->nl_current_all)
    return -1;
  for (int32_t l = 0; l < ndata; ++l) b[l] = 0.0;
  for (int32_t i = 0; i < m->nl_nonempty; ++i) {
    int32_t i_all = m->rowptr[i];
    if (i_all >= line_start && i_all <= line_end) {
      int32_t l = i_all - line_start;
      for (int64_t k = m->ijl[i] - 1; k < m->ijl[i + 1] - 1; ++k)
        b[l] = b[l] + (double)m->sa[k] * x[m->ija[k] - param_shift - 1];
    }
  }
  return 0;
}

/* sparse_matrix_add_trans_mult_vector, sparse_matrix.f90:388-405
 * (accumulation order into b(j) = stored-row order) */
void orc_csr_add_trans_mult_vector(const orc_csr *m, const double *x, double *b) {
  for (int32_t i = 0; i < m->nl_nonempty; ++i) {
    int32_t i_all = m->rowptr[i] - 1;
    for (int64_t k = m->ijl[i] - 1; k < m->ijl[i + 1] - 1; ++k) {
      int32_t j = m->ija[k] - 1;
      b[j] = b[j] + (double)m->sa[k] * x[i_all];
    }
  }
}

/* sparse_matrix_trans_mult_vector, sparse_matrix.f90:373-382 */
void orc_csr_trans_mult_vector(const orc_csr *m, const double *x, double *b) {
  for (int32_t j = 0; j < m->ncolumns; ++j) b[j] = 0.0;
  orc_csr_add_trans_mult_vector(m, x, b);
}

/* sparse_matrix_normalize_columns, sparse_matrix.f90:414-443 (test-only in the
 * reference; note it loops i = 1..nl, i.e. assumes no empty rows) */
void orc_csr_normalize_columns(orc_csr *m, double *column_norm) {
  for (int32_t j = 0; j < m->ncolumns; ++j) column_norm[j] = 0.0;
  for (int32_t i = 0; i < m->nl; ++i)
    for (int64_t k = m->ijl[i] - 1; k < m->ijl[i + 1] - 1; ++k) {
      /* sa(k)**2 is evaluated in real(4), then promoted for the sum */
      float sq = m->sa[k] * m->sa[k];
      column_norm[m->ija[k] - 1] = column_norm[m->ija[k] - 1] + (double)sq;
    }
  for (int32_t j = 0; j < m->ncolumns; ++j) column_norm[j] = sqrt(column_norm[j]);
  for (int32_t i = 0; i < m->nl; ++i)
    for (int64_t k = m->ijl[i] - 1; k < m->ijl[i + 1] - 1; ++k) {
      int32_t j = m->ija[k] - 1;
      if (column_norm[j] != 0.0) m->sa[k] = (float)((double)m->sa[k] / column_norm[j]);
    }
}

/* ========================================================================== */
/* wavelet_transform -- src/utils/wavelet_transform.F90                        */
/*                                                                            */
/* The reference sweeps whole (n-1)-D slabs `s(ig,:,:) = s(ig,:,:) - ...`;     */
/* every slab statement is elementwise, so the transform is an independent     */
/* 1-D multi-scale lifting on every line along the active axis. The oracle is  */
/* written line-wise; per element the operation sequence is identical.         */
/* ========================================================================== */

/* nscale = int(log(real(n))/log(2.)), wavelet_transform.F90:85 -- computed in
 * floating point exactly like the reference (equals floor(log2 n) for n<5000). */
static int orc_nscale(int n) { return (int)(log((double)n) / log(2.0)); }

/* pair geometry of one scale, wavelet_transform.F90:97-101 (0-based here):
 * low element of pair i at i*step, high element at step/2 + i*step, i < ng.   */
static int orc_npairs(int L, int step) {
  int ngmin = step / 2 + 1;
  int ngmax = ngmin + ((L - ngmin) / step) * step;
  return (ngmax - ngmin) / step + 1;
}

/* Haar3D one line, wavelet_transform.F90:96-150 */
static void haar_line_fwd(double *p, long st, int L) {
  const double sq2 = sqrt(2.0);
  int nscale = orc_nscale(L);
  for (int istep = 1; istep <= nscale; ++istep) {
    int step = 1 << istep, half = step / 2, ng = orc_npairs(L, step);
    for (int i = 0; i < ng; ++i) {                       /* predict  :103-116 */
      double *lo = p + (long)(i * step) * st, *hi = p + (long)(i * step + half) * st;
      *hi = *hi - *lo;
    }
    for (int i = 0; i < ng; ++i) {                       /* update   :118-131 */
      double *lo = p + (long)(i * step) * st, *hi = p + (long)(i * step + half) * st;
      *lo = *lo + *hi / 2.0;
    }
    for (int i = 0; i < ng; ++i) {                       /* normalise:133-149 */
      double *lo = p + (long)(i * step) * st, *hi = p + (long)(i * step + half) * st;
      *lo = *lo * sq2;
      *hi = *hi / sq2;
    }
  }
}

/* iHaar3D one line, wavelet_transform.F90:179-233 */
static void haar_line_inv(double *p, long st, int L) {
  const double sq2 = sqrt(2.0);
  int nscale = orc_nscale(L);
  for (int istep = nscale; istep >= 1; --istep) {
    int step = 1 << istep, half = step / 2, ng = orc_npairs(L, step);
    for (int i = 0; i < ng; ++i) {
      double *lo = p + (long)(i * step) * st, *hi = p + (long)(i * step + half) * st;
      *lo = *lo / sq2;
      *hi = *hi * sq2;
    }
    for (int i = 0; i < ng; ++i) {
      double *lo = p + (long)(i * step) * st, *hi = p + (long)(i * step + half) * st;
      *lo = *lo - *hi / 2.0;
    }
    for (int i = 0; i < ng; ++i) {
      double *lo = p + (long)(i * step) * st, *hi = p + (long)(i * step + half) * st;
      *hi = *hi + *lo;
    }
  }
}

/* DaubD43D one line, wavelet_transform.F90:271-364 */
static void d4_line_fwd(double *p, long st, int L) {
  const double c0 = sqrt(3.0), c1 = sqrt(3.0) / 4.0, c2 = (sqrt(3.0) - 2.0) / 4.0;
  const double c3 = (sqrt(3.0) - 1.0) / sqrt(2.0), c4 = (sqrt(3.0) + 1.0) / sqrt(2.0);
  int nscale = orc_nscale(L);
  for (int istep = 1; istep <= nscale; ++istep) {
    int step = 1 << istep, half = step / 2, ng = orc_npairs(L, step);
#define LO(i) p[(long)((i) * step) * st]
#define HI(i) p[(long)((i) * step + half) * st]
    for (int i = 0; i < ng; ++i) LO(i) = LO(i) + HI(i) * c0;            /* update 1 :280-293 */
    HI(0) = HI(0) - LO(0) * c1 - LO(ng - 1) * c2;                       /* predict, wrap :297-305 */
    for (int i = 1; i < ng; ++i) HI(i) = HI(i) - LO(i) * c1 - LO(i - 1) * c2;   /* :307-319 */
    for (int i = 0; i < ng - 1; ++i) LO(i) = LO(i) - HI(i + 1);         /* update 2 :321-334 */
    LO(ng - 1) = LO(ng - 1) - HI(0);                                    /* wrap :336-345 */
    for (int i = 0; i < ng; ++i) { LO(i) = LO(i) * c3; HI(i) = HI(i) * c4; }    /* :347-363 */
  }
}

/* iDaubD43D one line, wavelet_transform.F90:402-495 */
static void d4_line_inv(double *p, long st, int L) {
  const double c0 = sqrt(3.0), c1 = sqrt(3.0) / 4.0, c2 = (sqrt(3.0) - 2.0) / 4.0;
  const double c3 = (sqrt(3.0) - 1.0) / sqrt(2.0), c4 = (sqrt(3.0) + 1.0) / sqrt(2.0);
  int nscale = orc_nscale(L);
  for (int istep = nscale; istep >= 1; --istep) {
    int step = 1 << istep, half = step / 2, ng = orc_npairs(L, step);
    for (int i = 0; i < ng; ++i) { LO(i) = LO(i) * c4; HI(i) = HI(i) * c3; }    /* :411-427 */
    for (int i = ng - 2; i >= 0; --i) LO(i) = LO(i) + HI(i + 1);        /* update 2 :429-442 */
    LO(ng - 1) = LO(ng - 1) + HI(0);                                    /* :444-453 */
    for (int i = ng - 1; i >= 1; --i) HI(i) = HI(i) + LO(i) * c1 + LO(i - 1) * c2;   /* :455-468 */
    HI(0) = HI(0) + LO(0) * c1 + LO(ng - 1) * c2;                       /* :470-479 */
    for (int i = 0; i < ng; ++i) LO(i) = LO(i) - HI(i) * c0;            /* update 1 :481-494 */
#undef LO
#undef HI
  }
}

typedef void (*line_fn)(double *, long, int);

/* axis order 1,2,3 with all scales of an axis before the next axis
 * (wavelet_transform.F90:83-151). s is Fortran order s(n1,n2,n3). */
static void apply_3d(double *s, int n1, int n2, int n3, line_fn fn) {
  long s1 = 1, s2 = n1, s3 = (long)n1 * n2;
  for (int k = 0; k < n3; ++k)
    for (int j = 0; j < n2; ++j) fn(s + j * s2 + k * s3, s1, n1);
  for (int k = 0; k < n3; ++k)
    for (int i = 0; i < n1; ++i) fn(s + i * s1 + k * s3, s2, n2);
  for (int j = 0; j < n2; ++j)
    for (int i = 0; i < n1; ++i) fn(s + i * s1 + j * s2, s3, n3);
}

void orc_haar3d(double *s, int n1, int n2, int n3)    { apply_3d(s, n1, n2, n3, haar_line_fwd); }
void orc_ihaar3d(double *s, int n1, int n2, int n3)   { apply_3d(s, n1, n2, n3, haar_line_inv); }
void orc_daubd43d(double *s, int n1, int n2, int n3)  { apply_3d(s, n1, n2, n3, d4_line_fwd); }
void orc_idaubd43d(double *s, int n1, int n2, int n3) { apply_3d(s, n1, n2, n3, d4_line_inv); }

/* forward_wavelet / inverse_wavelet, wavelet_transform.F90:37-70 */
int orc_forward_wavelet(double *s, int n1, int n2, int n3, int wavelet_type) {
  if (wavelet_type == 1) orc_haar3d(s, n1, n2, n3);
  else if (wavelet_type == 2) orc_daubd43d(s, n1, n2, n3);
  else return -1;
  return 0;
}
int orc_inverse_wavelet(double *s, int n1, int n2, int n3, int wavelet_type) {
  if (wavelet_type == 1) orc_ihaar3d(s, n1, n2, n3);
  else if (wavelet_type == 2) orc_idaubd43d(s, n1, n2, n3);
  else return -1;
  return 0;
}

/* ========================================================================== */
/* lsqr_solver -- src/inversion/lsqr_solver2.F90                               */
/* ========================================================================== */

/* Fortran norm2 intrinsic (scaled sum of squares, as libgfortran does). */
double orc_norm2(int64_t n, const double *x) {
  double scale = 1.0, ssq = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    if (x[i] != 0.0) {
      double absx = fabs(x[i]);
      if (scale < absx) { double v = scale / absx; ssq = 1.0 + ssq * (v * v); scale = absx; }
      else { double v = absx / scale; ssq = ssq + v * v; }
    }
  }
  return scale * sqrt(ssq);
}

/* normalize, lsqr_solver2.F90:501-530 (single rank: the Allreduce is identity) */
static int orc_normalize(int64_t n, double *x, double *s, int in_parallel) {
  if (in_parallel) {
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) acc += x[i] * x[i];
    *s = sqrt(acc);
  } else {
    *s = orc_norm2(n, x);
  }
  if (*s == 0.0) return -1;
  double ss = 1.0 / *s;
  for (int64_t i = 0; i < n; ++i) x[i] = ss * x[i];
  return 0;
}

/* apply_soft_thresholding, lsqr_solver2.F90:478-494 */
static void orc_soft_threshold(double *x, int64_t n, double gamma) {
  for (int64_t i = 0; i < n; ++i) {
    if (fabs(x[i]) <= gamma) x[i] = 0.0;
    else if (x[i] <= -gamma) x[i] = x[i] + gamma;
    else if (x[i] >= gamma) x[i] = x[i] - gamma;
  }
}

/* lsqr_solve, lsqr_solver2.F90:321-473 */
int orc_lsqr_solve(int32_t nlines, int32_t nelements, int32_t niter, double rmin, double gamma,
                   const orc_csr *matrix, double *u, double *x, double *r_hist, int32_t *iters) {
  if (iters) *iters = 0;
  if (matrix->nl != nlines || matrix->ncolumns != nelements) return -1;        /* :342-345 */
  double *v = (double *)calloc((size_t)nelements, sizeof(double));
  double *w = (double *)calloc((size_t)nelements, sizeof(double));
  for (int32_t i = 0; i < nelements; ++i) x[i] = 0.0;                          /* :352 */
  int rc = 0;
  double alpha, beta, rho, rhobar, phi, phibar, theta, b1, c, r, s, t1, t2, rho_inv;
  if (orc_norm2(nlines, u) == 0.0) goto done;                                  /* :355-358 */
  if (orc_normalize(nlines, u, &beta, 0) != 0) { rc = -2; goto done; }         /* :361-364 */
  b1 = beta;
  orc_csr_trans_mult_vector(matrix, u, v);                                     /* :369 */
  if (orc_normalize(nelements, v, &alpha, 1) != 0) { rc = -3; goto done; }     /* :372-375 */
  rhobar = alpha; phibar = beta;
  memcpy(w, v, sizeof(double) * (size_t)nelements);
  int32_t iter = 1;
  r = 1.0;
  while (iter <= niter && r > rmin) {                                          /* :385 */
    for (int32_t i = 0; i < nlines; ++i) u[i] = -alpha * u[i];                 /* :390-394 */
    orc_csr_add_mult_vector(matrix, v, u);                                     /* :397 */
    orc_normalize(nlines, u, &beta, 0);                                        /* :404-408 */
    for (int32_t i = 0; i < nelements; ++i) v[i] = -beta * v[i];               /* :411 */
    orc_csr_add_trans_mult_vector(matrix, u, v);                               /* :414 */
    orc_normalize(nelements, v, &alpha, 1);                                    /* :417-421 */
    rho = sqrt(rhobar * rhobar + beta * beta);                                 /* :424 */
    if (rho == 0.0) break;                                                     /* :427-430 */
    rho_inv = 1.0 / rho;
    c = rhobar * rho_inv; s = beta * rho_inv; theta = s * alpha; rhobar = -c * alpha;
    phi = c * phibar; phibar = s * phibar; t1 = phi * rho_inv; t2 = -theta * rho_inv;
    for (int32_t i = 0; i < nelements; ++i) x[i] = t1 * w[i] + x[i];           /* :445 */
    for (int32_t i = 0; i < nelements; ++i) w[i] = t2 * w[i] + v[i];           /* :446 */
    if (gamma != 0.0) orc_soft_threshold(x, nelements, gamma);                 /* :448-451 */
    r = phibar / b1;
    if (r_hist) r_hist[iter - 1] = r;
    if (fabs(rhobar) < 1.e-30f) { iter += 1; break; }   /* :460-463 exits BEFORE iter++; count it as done */
    iter += 1;
  }
  if (iters) *iters = iter - 1;
done:
  free(v); free(w);
  return rc;
}

/* apply_wavelet_transform for nbproc == 1, src/inversion/wavelet_utils.F90:37-72:
 * v has shape (nelements, ncomponents, nproblems); each active problem and
 * component is transformed as a separate nx*ny*nz volume. */
static void orc_apply_wavelet(int32_t nelements, int nx, int ny, int nz, int ncomp, double *v, int fwd,
                              int type, int nproblems, const int32_t *solve_problem) {
  for (int i = 0; i < nproblems; ++i) {
    if (!solve_problem[i]) continue;
    for (int k = 0; k < ncomp; ++k) {
      double *p = v + ((size_t)i * ncomp + k) * (size_t)nelements;
      if (fwd) orc_forward_wavelet(p, nx, ny, nz, type);
      else orc_inverse_wavelet(p, nx, ny, nz, type);
    }
  }
}

/* lsqr_solve_sensit, lsqr_solver2.F90:47-308 (single rank) */
int orc_lsqr_solve_sensit(int32_t nlines, int32_t ncolumns, int32_t niter, double rmin, double gamma,
                          double target_misfit, const orc_csr *S, const orc_csr *C,
                          double *u, double *x, const int32_t solve_problem[2], int32_t nelements,
                          int32_t nx, int32_t ny, int32_t nz, int32_t ncomponents,
                          int32_t compression_type, int32_t wavelet_domain,
                          double *r_hist, int32_t *iters) {
  if (iters) *iters = 0;
  if (S->nl + C->nl != nlines || S->ncolumns != ncolumns || C->ncolumns != ncolumns) return -1;  /* :85-89 */
  int32_t nls = S->nl;
  int calc_misfit = target_misfit > 0.0;
  int wav = (compression_type > 0 && !wavelet_domain);
  double *v = (double *)calloc((size_t)ncolumns, sizeof(double));
  double *w = (double *)calloc((size_t)ncolumns, sizeof(double));
  double *v2 = (double *)calloc((size_t)ncolumns, sizeof(double));
  double *b0 = NULL, *Sx = NULL;
  if (calc_misfit) {
    b0 = (double *)calloc((size_t)nls, sizeof(double));
    Sx = (double *)calloc((size_t)nls, sizeof(double));
    memcpy(b0, u, sizeof(double) * (size_t)nls);                               /* :110 */
  }
  for (int32_t i = 0; i < ncolumns; ++i) x[i] = 0.0;                           /* :120 */
  int rc = 0;
  double alpha, beta, rho, rhobar, phi, phibar, theta, b1, c, r, s, t1, t2, rho_inv, misfit;
  if (orc_norm2(nlines, u) == 0.0) goto done;                                  /* :123-126 */
  if (orc_normalize(nlines, u, &beta, 0) != 0) { rc = -2; goto done; }         /* :129-132 */
  b1 = beta;
  orc_csr_trans_mult_vector(S, u, v2);                                         /* :137 */
  if (wav) orc_apply_wavelet(nelements, nx, ny, nz, ncomponents, v2, 0, compression_type, 2, solve_problem);
  memcpy(v, v2, sizeof(double) * (size_t)ncolumns);                            /* :145 */
  orc_csr_add_trans_mult_vector(C, u + nls, v);                                /* :147 */
  if (orc_normalize(ncolumns, v, &alpha, 1) != 0) { rc = -3; goto done; }      /* :150-153 */
  rhobar = alpha; phibar = beta;
  memcpy(w, v, sizeof(double) * (size_t)ncolumns);
  int32_t iter = 1;
  r = 1.0;
  while (iter <= niter && r > rmin) {                                          /* :163 */
    if (calc_misfit) {                                                         /* :168-189 */
      memcpy(v2, x, sizeof(double) * (size_t)ncolumns);
      if (wav) orc_apply_wavelet(nelements, nx, ny, nz, ncomponents, v2, 1, compression_type, 2, solve_problem);
      orc_csr_mult_vector(S, v2, Sx);
      double acc = 0.0;
      for (int32_t i = 0; i < nls; ++i) acc += (Sx[i] - b0[i]) * (Sx[i] - b0[i]);
      misfit = sqrt(acc / (double)nls);
      if (misfit <= target_misfit) break;
    }
    for (int32_t i = 0; i < nlines; ++i) u[i] = -alpha * u[i];                 /* :194-198 */
    memcpy(v2, v, sizeof(double) * (size_t)ncolumns);                          /* :200 */
    if (wav) orc_apply_wavelet(nelements, nx, ny, nz, ncomponents, v2, 1, compression_type, 2, solve_problem);
    orc_csr_add_mult_vector(S, v2, u);                                         /* :209 */
    orc_csr_add_mult_vector(C, v, u + nls);                                    /* :211 */
    orc_normalize(nlines, u, &beta, 0);                                        /* :218-222 */
    for (int32_t i = 0; i < ncolumns; ++i) v[i] = -beta * v[i];                /* :225 */
    orc_csr_trans_mult_vector(S, u, v2);                                       /* :228 */
    if (wav) orc_apply_wavelet(nelements, nx, ny, nz, ncomponents, v2, 0, compression_type, 2, solve_problem);
    for (int32_t i = 0; i < ncolumns; ++i) v[i] = v[i] + v2[i];                /* :236 */
    orc_csr_add_trans_mult_vector(C, u + nls, v);                              /* :238 */
    orc_normalize(ncolumns, v, &alpha, 1);                                     /* :241-245 */
    rho = sqrt(rhobar * rhobar + beta * beta);                                 /* :248 */
    if (rho == 0.0) break;                                                     /* :251-254 */
    rho_inv = 1.0 / rho;
    c = rhobar * rho_inv; s = beta * rho_inv; theta = s * alpha; rhobar = -c * alpha;
    phi = c * phibar; phibar = s * phibar; t1 = phi * rho_inv; t2 = -theta * rho_inv;
    for (int32_t i = 0; i < ncolumns; ++i) x[i] = t1 * w[i] + x[i];            /* :269 */
    for (int32_t i = 0; i < ncolumns; ++i) w[i] = t2 * w[i] + v[i];            /* :270 */
    if (gamma != 0.0) orc_soft_threshold(x, ncolumns, gamma);                  /* :272-275 */
    r = phibar / b1;                                                           /* :277-281 */
    if (r_hist) r_hist[iter - 1] = r;
    iter += 1;                                                                 /* :283 */
    if (fabs(rhobar) < 1.e-30f) break;                                         /* :286-289 */
  }
  if (iters) *iters = iter - 1;
done:
  free(v); free(w); free(v2); free(b0); free(Sx);
  return rc;
}

/* ========================================================================== */
/* gravity_field -- src/forward/gravmag/grav/gravity_field.f90                 */
/* ========================================================================== */

/* G_grav is declared with a single-precision literal (gravity_field.f90:26),
 * so the value actually used is float(6.674e-11) promoted to double. */
static const double ORC_G_GRAV = (double)6.674e-11f;
/* PI, global_typedefs.F90:52 */
static const double ORC_PI = 3.1415926535897932385;

/* graviprism_z, gravity_field.f90:131-195 */
int orc_graviprism_z(int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                     const double *Z1, const double *Z2, double xd, double yd, double zd, double *lineZ) {
  const double twopi = 2.0 * ORC_PI;
  const double signo[2] = {-1.0, 1.0};
  for (int32_t i = 0; i < n; ++i) {
    double XX[2] = {xd - X1[i], xd - X2[i]};
    double YY[2] = {yd - Y1[i], yd - Y2[i]};
    double ZZ[2] = {zd - Z1[i], zd - Z2[i]};
    double gz = 0.0;
    for (int K = 0; K < 2; ++K)
      for (int L = 0; L < 2; ++L)
        for (int M = 0; M < 2; ++M) {
          double dmu = signo[K] * signo[L] * signo[M];
          double Rs = sqrt(XX[K] * XX[K] + YY[L] * YY[L] + ZZ[M] * ZZ[M]);
          double arg3 = atan2(XX[K] * YY[L], ZZ[M] * Rs);
          if (arg3 < 0) arg3 = arg3 + twopi;
          double arg4 = Rs + XX[K];
          double arg5 = Rs + YY[L];
          if (arg4 <= 0.) return 1;
          if (arg5 <= 0.) return 2;
          arg4 = log(arg4);
          arg5 = log(arg5);
          gz = gz + dmu * (ZZ[M] * arg3 - XX[K] * arg5 - YY[L] * arg4);
        }
    lineZ[i] = ORC_G_GRAV * gz;
  }
  return 0;
}

/* gradiprism_zz, gravity_field.f90:314-364 */
void orc_gradiprism_zz(int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                       const double *Z1, const double *Z2, double xd, double yd, double zd, double *lineZZ) {
  const double twopi = 2.0 * ORC_PI;
  const double signo[2] = {-1.0, 1.0};
  for (int32_t i = 0; i < n; ++i) {
    double XX[2] = {xd - X1[i], xd - X2[i]};
    double YY[2] = {yd - Y1[i], yd - Y2[i]};
    double ZZ[2] = {-(zd - Z1[i]), -(zd - Z2[i])};
    double gzz = 0.0;
    for (int K = 0; K < 2; ++K)
      for (int L = 0; L < 2; ++L)
        for (int M = 0; M < 2; ++M) {
          double dmu = signo[K] * signo[L] * signo[M];
          double Rs = sqrt(XX[K] * XX[K] + YY[L] * YY[L] + ZZ[M] * ZZ[M]);
          double vzz = -atan2(XX[K] * YY[L], Rs * ZZ[M]);
          if (vzz < 0) vzz = vzz + twopi;
          gzz = gzz + dmu * vzz;
        }
    lineZZ[i] = ORC_G_GRAV * gzz;
  }
}

/* gradiprism_full, gravity_field.f90:207-309: the six components of the gravity gradient tensor of every prism at one
 * station. lines[d*n + i], d = 0..5 in the order the caller stores them (sensitivity_gravmag.F90:207-209):
 * XX, YY, ZZ, XY, YZ, ZX. Returns 3 for "Zero denominator in gradiprism_full!", 4 for "Bad log argument in
 * gradiprism_full!" (:271-280, exit_MPI in the reference), 0 otherwise. */
int orc_gradiprism_full(int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                        const double *Z1, const double *Z2, double xd, double yd, double zd, double *lines) {
  const double twopi = 2.0 * ORC_PI;
  const double signo[2] = {-1.0, 1.0};
  for (int32_t i = 0; i < n; ++i) {
    double XX[2] = {xd - X1[i], xd - X2[i]};
    double YY[2] = {yd - Y1[i], yd - Y2[i]};
    double ZZ[2] = {-(zd - Z1[i]), -(zd - Z2[i])};
    double gxx = 0.0, gxy = 0.0, gyy = 0.0, gzx = 0.0, gyz = 0.0, gzz = 0.0;
    for (int K = 0; K < 2; ++K)
      for (int L = 0; L < 2; ++L)
        for (int M = 0; M < 2; ++M) {
          double dmu = signo[K] * signo[L] * signo[M];
          double Rs = sqrt(XX[K] * XX[K] + YY[L] * YY[L] + ZZ[M] * ZZ[M]);
          double vxx = atan2(XX[K] * YY[L], XX[K] * XX[K] + Rs * ZZ[M] + ZZ[M] * ZZ[M]);
          double vyy = atan2(XX[K] * YY[L], Rs * Rs + Rs * ZZ[M] - XX[K] * XX[K]);
          double vzz = -atan2(XX[K] * YY[L], Rs * ZZ[M]);
          if (vxx < 0) vxx = vxx + twopi;
          if (vyy < 0) vyy = vyy + twopi;
          if (vzz < 0) vzz = vzz + twopi;
          double arg1 = Rs + ZZ[M];
          double arg21 = Rs - YY[L], arg22 = Rs + YY[L];
          double arg31 = Rs - XX[K], arg32 = Rs + XX[K];
          if (arg22 == 0. || arg32 == 0.) return 3;
          double arg2 = arg21 / arg22;
          double arg3 = arg31 / arg32;
          if (arg1 <= 0. || arg2 <= 0. || arg3 <= 0.) return 4;
          double vxy = log(arg1);
          double vzx = 0.5 * log(arg2);
          double vyz = 0.5 * log(arg3);
          gxx = gxx + dmu * vxx;
          gyy = gyy + dmu * vyy;
          gzz = gzz + dmu * vzz;
          gxy = gxy + dmu * vxy;
          gyz = gyz + dmu * vyz;
          gzx = gzx + dmu * vzx;
        }
    lines[0 * (size_t)n + i] = ORC_G_GRAV * gxx;
    lines[1 * (size_t)n + i] = ORC_G_GRAV * gyy;
    lines[2 * (size_t)n + i] = ORC_G_GRAV * gzz;
    lines[3 * (size_t)n + i] = ORC_G_GRAV * gxy;
    lines[4 * (size_t)n + i] = ORC_G_GRAV * gyz;
    lines[5 * (size_t)n + i] = ORC_G_GRAV * gzx;
  }
  return 0;
}

/* ========================================================================== */
/* magnetic_field -- src/forward/gravmag/mag/magnetic_field.f90                */
/* ========================================================================== */

/* dircos, magnetic_field.f90:91-110 */
void orc_dircos(double incl, double decl, double azim, double *a, double *b, double *c) {
  const double d2rad = ORC_PI / 180.0;
  double decl2 = fmod(450.0 - decl, 360.0);
  double xincl = incl * d2rad, xdecl = decl2 * d2rad, xazim = azim * d2rad;
  *a = cos(xincl) * cos(xdecl - xazim);
  *b = cos(xincl) * sin(xdecl - xazim);
  *c = sin(xincl);
}

/* sharmbox, magnetic_field.f90:321-457 (eps = 0). Returns 1/2 on the
 * "grid boundary coincides with data position" stops. */
int orc_sharmbox(double x0, double y0, double z0, double x1, double y1, double z1,
                 double x2, double y2, double z2, double tsx[3], double tsy[3], double tsz[3]) {
  const double eps = 0.;
  double rx1 = x1 - x0 + eps, rx2 = x2 - x0 + eps;
  double ry1 = y1 - y0 + eps, ry2 = y2 - y0 + eps;
  double rz1 = z1 - z0 + eps, rz2 = z2 - z0 + eps;
  if (rx1 == 0. || rx2 == 0.) return 1;
  if (ry1 == 0. || ry2 == 0.) return 2;
  double rx1sq = rx1 * rx1, rx2sq = rx2 * rx2, ry1sq = ry1 * ry1, ry2sq = ry2 * ry2;
  double rz1sq = rz1 * rz1, rz2sq = rz2 * rz2;
  double R1 = ry2sq + rx2sq, R2 = ry2sq + rx1sq, R3 = ry1sq + rx2sq, R4 = ry1sq + rx1sq;
  double arg1 = sqrt(rz2sq + R2), arg2 = sqrt(rz2sq + R1), arg3 = sqrt(rz1sq + R1), arg4 = sqrt(rz1sq + R2);
  double arg5 = sqrt(rz2sq + R3), arg6 = sqrt(rz2sq + R4), arg7 = sqrt(rz1sq + R4), arg8 = sqrt(rz1sq + R3);
  tsx[0] = atan2(ry1 * rz2, (rx2 * arg5 + eps)) - atan2(ry2 * rz2, (rx2 * arg2 + eps)) +
           atan2(ry2 * rz1, (rx2 * arg3 + eps)) - atan2(ry1 * rz1, (rx2 * arg8 + eps)) +
           atan2(ry2 * rz2, (rx1 * arg1 + eps)) - atan2(ry1 * rz2, (rx1 * arg6 + eps)) +
           atan2(ry1 * rz1, (rx1 * arg7 + eps)) - atan2(ry2 * rz1, (rx1 * arg4 + eps));
  tsy[0] = log((rz2 + arg2 + eps) / (rz1 + arg3 + eps)) - log((rz2 + arg1 + eps) / (rz1 + arg4 + eps)) +
           log((rz2 + arg6 + eps) / (rz1 + arg7 + eps)) - log((rz2 + arg5 + eps) / (rz1 + arg8 + eps));
  tsy[1] = atan2(rx1 * rz2, (ry2 * arg1 + eps)) - atan2(rx2 * rz2, (ry2 * arg2 + eps)) +
           atan2(rx2 * rz1, (ry2 * arg3 + eps)) - atan2(rx1 * rz1, (ry2 * arg4 + eps)) +
           atan2(rx2 * rz2, (ry1 * arg5 + eps)) - atan2(rx1 * rz2, (ry1 * arg6 + eps)) +
           atan2(rx1 * rz1, (ry1 * arg7 + eps)) - atan2(rx2 * rz1, (ry1 * arg8 + eps));
  R1 = ry2sq + rz1sq; R2 = ry2sq + rz2sq; R3 = ry1sq + rz1sq; R4 = ry1sq + rz2sq;
  arg1 = sqrt(rx1sq + R1); arg2 = sqrt(rx2sq + R1); arg3 = sqrt(rx1sq + R2); arg4 = sqrt(rx2sq + R2);
  arg5 = sqrt(rx1sq + R3); arg6 = sqrt(rx2sq + R3); arg7 = sqrt(rx1sq + R4); arg8 = sqrt(rx2sq + R4);
  tsy[2] = log((rx1 + arg1 + eps) / (rx2 + arg2 + eps)) - log((rx1 + arg3 + eps) / (rx2 + arg4 + eps)) +
           log((rx1 + arg7 + eps) / (rx2 + arg8 + eps)) - log((rx1 + arg5 + eps) / (rx2 + arg6 + eps));
  R1 = rx2sq + rz1sq; R2 = rx2sq + rz2sq; R3 = rx1sq + rz1sq; R4 = rx1sq + rz2sq;
  arg1 = sqrt(ry1sq + R1); arg2 = sqrt(ry2sq + R1); arg3 = sqrt(ry1sq + R2); arg4 = sqrt(ry2sq + R2);
  arg5 = sqrt(ry1sq + R3); arg6 = sqrt(ry2sq + R3); arg7 = sqrt(ry1sq + R4); arg8 = sqrt(ry2sq + R4);
  tsx[2] = log((ry1 + arg1 + eps) / (ry2 + arg2 + eps)) - log((ry1 + arg3 + eps) / (ry2 + arg4 + eps)) +
           log((ry1 + arg7 + eps) / (ry2 + arg8 + eps)) - log((ry1 + arg5 + eps) / (ry2 + arg6 + eps));
  tsz[2] = -1 * (tsx[0] + tsy[1]);
  tsz[1] = tsy[2];
  tsx[1] = tsy[0];
  tsz[0] = tsx[2];
  return 0;
}

static double min2(double a, double b) { return a < b ? a : b; }

/* magprism, magnetic_field.f90:118-297. sensit_line is Fortran-ordered
 * (nelements, nmodel_comp, ndata_comp). */
int orc_magprism(int32_t n, int32_t nmc, int32_t ndc,
                 const double *X1, const double *X2, const double *Y1, const double *Y2,
                 const double *Z1, const double *Z2, double xd, double yd, double zd,
                 double mi, double md, double theta, double intensity, double *sl) {
  const double mu0 = 4.0 * ORC_PI * 1.e-7, T2nT = 1.e+9;
  double magv[3];
  orc_dircos(mi, md, theta, &magv[0], &magv[1], &magv[2]);                     /* :64-77 */
  if (!((nmc == 1 || nmc == 3) && (ndc == 1 || ndc == 3))) return -1;
#define SL(i, k, d) sl[(size_t)(i) + (size_t)n * ((size_t)(k) + (size_t)nmc * (size_t)(d))]
  for (int32_t i = 0; i < n; ++i) {
    double tx[3], ty[3], tz[3];
    if ((X1[i] < xd) && (X2[i] > xd) && (Y1[i] < yd) && (Y2[i] > yd) && (Z1[i] < zd) && (Z2[i] > zd)) {
      /* observation point inside the cell: 6 sub-voxels around a void, :139-224.
       * `width = 0.1` is a single-precision literal in the reference. */
      double width = (double)0.1f;
      double min_clr = min2(min2(min2(fabs(xd - X1[i]), fabs(xd - X2[i])),
                                 min2(fabs(yd - Y1[i]), fabs(yd - Y2[i]))),
                            min2(fabs(zd - Z1[i]), fabs(zd - Z2[i])));
      if (width > min_clr) width = 0.5 * min_clr;
      double bx1[6] = {X1[i], X1[i], X1[i], xd + width, xd - width, xd - width};
      double bx2[6] = {X2[i], X2[i], xd - width, X2[i], xd + width, xd + width};
      double by1[6] = {Y1[i], Y1[i], Y1[i], Y1[i], Y1[i], yd + width};
      double by2[6] = {Y2[i], Y2[i], Y2[i], Y2[i], yd - width, Y2[i]};
      double bz1[6] = {Z1[i], zd + width, zd - width, zd - width, zd - width, zd - width};
      double bz2[6] = {zd - width, Z2[i], zd + width, zd + width, zd + width, zd + width};
      for (int q = 0; q < 3; ++q) tx[q] = ty[q] = tz[q] = 0.0;
      for (int j = 0; j < 6; ++j) {
        double ax[3], ay[3], az[3];
        int rc = orc_sharmbox(xd, yd, zd, bx1[j], by1[j], bz1[j], bx2[j], by2[j], bz2[j], ax, ay, az);
        if (rc) return rc;
        for (int q = 0; q < 3; ++q) { tx[q] = tx[q] + ax[q]; ty[q] = ty[q] + ay[q]; tz[q] = tz[q] + az[q]; }
      }
    } else {
      int rc = orc_sharmbox(xd, yd, zd, X1[i], Y1[i], Z1[i], X2[i], Y2[i], Z2[i], tx, ty, tz);
      if (rc) return rc;
    }
    if (nmc == 1) {                                                             /* :240-258 */
      double mx = tx[0] * magv[0] + tx[1] * magv[1] + tx[2] * magv[2];
      double my = ty[0] * magv[0] + ty[1] * magv[1] + ty[2] * magv[2];
      double mz = tz[0] * magv[0] + tz[1] * magv[1] + tz[2] * magv[2];
      if (ndc == 1) SL(i, 0, 0) = mx * magv[0] + my * magv[1] + mz * magv[2];
      else { SL(i, 0, 0) = mx; SL(i, 0, 1) = my; SL(i, 0, 2) = mz; }
    } else {                                                                    /* :260-278 */
      for (int k = 0; k < 3; ++k) {
        if (ndc == 1) SL(i, k, 0) = tx[k] * magv[0] + ty[k] * magv[1] + tz[k] * magv[2];
        else { SL(i, k, 0) = tx[k]; SL(i, k, 1) = ty[k]; SL(i, k, 2) = tz[k]; }
      }
    }
  }
#undef SL
  size_t tot = (size_t)n * nmc * ndc;
  double mult = (nmc == 1) ? intensity : (mu0 * T2nT);                          /* :286-292 */
  for (size_t q = 0; q < tot; ++q) sl[q] = mult * sl[q];
  for (size_t q = 0; q < tot; ++q) sl[q] = sl[q] / (4.0 * ORC_PI);              /* :295 */
  return 0;
}

/* ========================================================================== */
/* row compression -- src/forward/gravmag/sensitivity_gravmag.F90:230-295      */
/* ========================================================================== */
static int cmp_double(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}

int32_t orc_compress_row(double *line, int32_t nx, int32_t ny, int32_t nz, int32_t compression_type,
                         int32_t nel_compressed, int32_t *cols, float *vals,
                         double *threshold_out, double *cost_full_out, double *cost_discarded_out) {
  int32_t N = nx * ny * nz, nel = 0;
  double threshold = 0.0, cost_full = 0.0, cost_discarded = 0.0;
  if (compression_type > 0) {
    for (int32_t p = 0; p < N; ++p) cost_full += line[p] * line[p];            /* :234 */
    orc_forward_wavelet(line, nx, ny, nz, compression_type);                   /* :237 */
    if (nel_compressed >= N) {
      threshold = -1.0;                                                        /* :244-246 */
    } else {
      double *sorted = (double *)malloc(sizeof(double) * (size_t)N);           /* :240-241 */
      for (int32_t p = 0; p < N; ++p) sorted[p] = fabs(line[p]);
      qsort(sorted, (size_t)N, sizeof(double), cmp_double);
      threshold = fabs(sorted[N - nel_compressed - 1]);                        /* :248-249 */
      free(sorted);
    }
    if (threshold < 1.e-30) threshold = 1.e-30;                                /* :252-256 */
    for (int32_t p = 0; p < N; ++p) {                                          /* :258-272 */
      if (fabs(line[p]) > threshold) {
        cols[nel] = p + 1;
        vals[nel] = (float)line[p];
        nel += 1;
      } else {
        cost_discarded += line[p] * line[p];
      }
    }
  } else {                                                                     /* :287-295 */
    nel = N;
    for (int32_t p = 0; p < N; ++p) { cols[p] = p + 1; vals[p] = (float)line[p]; }
  }
  if (threshold_out) *threshold_out = threshold;
  if (cost_full_out) *cost_full_out = cost_full;
  if (cost_discarded_out) *cost_discarded_out = cost_discarded;
  return nel;
}

/* ========================================================================== */
/* depth weight -- src/forward/gravmag/weights_gravmag.f90:46-250              */
/* ========================================================================== */
int orc_depth_weight(int32_t type, int32_t n, const double *X1, const double *X2, const double *Y1,
                     const double *Y2, const double *Z1, const double *Z2, int32_t ndata,
                     const double *xd, const double *yd, const double *zd,
                     double power, double beta, double Z0, double *cw) {
  if (type == 1) {                                                             /* :71-79, :204-223 */
    for (int32_t p = 0; p < n; ++p) {
      double depth = 0.5 * (Z1[p] + Z2[p]);
      if (!(depth + Z0 > 0.0)) return -1;
      cw[p] = pow(depth + Z0, -power / 2.0);
    }
  } else if (type == 2) {                                                      /* :81-138 */
    const double R0 = 0.1, dfactor = 0.25;
    for (int32_t p = 0; p < n; ++p) {
      double dVj = fabs((X2[p] - X1[p]) * (Y2[p] - Y1[p]) * (Z2[p] - Z1[p]));
      double dhx = dfactor * fabs(X2[p] - X1[p]);
      double dhy = dfactor * fabs(Y2[p] - Y1[p]);
      double dhz = dfactor * fabs(Z2[p] - Z1[p]);
      double wr = 0.0;
      for (int32_t j = 0; j < ndata; ++j) {
        double dx[2], dy[2], dz[2];
        dx[0] = pow(X1[p] + dhx - xd[j], 2.0); dy[0] = pow(Y1[p] + dhy - yd[j], 2.0);
        dz[0] = pow(Z1[p] + dhz - zd[j], 2.0);
        dx[1] = pow(X2[p] - dhx - xd[j], 2.0); dy[1] = pow(Y2[p] - dhy - yd[j], 2.0);
        dz[1] = pow(Z2[p] - dhz - zd[j], 2.0);
        double integral = 0.0;
        for (int ii = 0; ii < 2; ++ii)
          for (int jj = 0; jj < 2; ++jj)
            for (int kk = 0; kk < 2; ++kk) {
              double R = sqrt(dx[ii] + dy[jj] + dz[kk]);
              integral = integral + 1.0 / pow(R + R0, power);
            }
        integral = integral * dVj / 8.0;
        wr = wr + pow(integral, 2.0);
      }
      cw[p] = (1.0 / sqrt(dVj)) * pow(wr, beta / 4.0);
    }
  } else if (type == 3) {                                                      /* :140-161 */
    const double R0 = 0.01;
    for (int32_t p = 0; p < n; ++p) {
      double mindist = 1.e30;
      double cx = 0.5 * (X1[p] + X2[p]), cy = 0.5 * (Y1[p] + Y2[p]), cz = 0.5 * (Z1[p] + Z2[p]);
      for (int32_t j = 0; j < ndata; ++j) {
        double dist = sqrt(pow(cx - xd[j], 2.0) + pow(cy - yd[j], 2.0) + pow(cz - zd[j], 2.0));
        if (dist < mindist) mindist = dist;
      }
      cw[p] = sqrt(1.0 / pow(mindist + R0, power));
    }
  } else {
    return -2;
  }
  double norm = -1.e300;
  for (int32_t p = 0; p < n; ++p) {                                            /* :170-175 */
    cw[p] = cw[p] * sqrt(fabs((X2[p] - X1[p]) * (Y2[p] - Y1[p]) * (Z2[p] - Z1[p])));
    if (cw[p] > norm) norm = cw[p];
  }
  if (norm == 0.0) return -3;                                                  /* :228-250 */
  for (int32_t p = 0; p < n; ++p) cw[p] = cw[p] / norm;
  for (int32_t p = 0; p < n; ++p) {                                            /* :189-195 */
    if (cw[p] == 0.0) return -4;
    cw[p] = 1.0 / cw[p];
  }
  return 0;
}

/* ========================================================================== */
/* ADMM -- src/inversion/admm_method.F90:70-134                                */
/* xmin/xmax are Fortran-ordered (nlithos, n).                                 */
/* ========================================================================== */
void orc_admm_iterate(int32_t n, int32_t nlithos, const double *xmin, const double *xmax,
                      const double *x, double *z, double *u, double *x0) {
  for (int32_t i = 0; i < n; ++i) {
    double arg = x[i] + u[i];
    int inside = 0;
    for (int32_t j = 0; j < nlithos; ++j) {
      if (xmin[j + (size_t)nlithos * i] <= arg && arg <= xmax[j + (size_t)nlithos * i]) {
        inside = 1; z[i] = arg; break;
      }
    }
    if (!inside) {
      double mindist = 1.e30, closest = 0.0;
      for (int32_t j = 0; j < nlithos; ++j) {
        double val = fabs(xmin[j + (size_t)nlithos * i] - arg);
        if (val < mindist) { mindist = val; closest = xmin[j + (size_t)nlithos * i]; }
        val = fabs(xmax[j + (size_t)nlithos * i] - arg);
        if (val < mindist) { mindist = val; closest = xmax[j + (size_t)nlithos * i]; }
      }
      z[i] = closest;
    }
  }
  for (int32_t i = 0; i < n; ++i) u[i] = u[i] + x[i] - z[i];
  for (int32_t i = 0; i < n; ++i) x0[i] = z[i] - u[i];
}

/* ========================================================================== */
/* Constraint-matrix producers.                             */
/* Serial restatements with the rank's column slab given explicitly:           */
/* nsmaller = get_nsmaller(nelements, myrank, nbproc), local cells             */
/* nsmaller+1 .. nsmaller+nelements.  b_RHS is the constraint part of the      */
/* right-hand side (b_RHS(lc:) in joint_inverse_problem.F90:465).              */
/* ========================================================================== */

/* damping_add + damping_add_RHS + get_norm_multiplier, src/inversion/damping.F90:97-261.
 * model / model_ref / column_weight / local_weight (may be NULL) hold the FULL model here (the serial
 * union of all ranks' slabs); only entries of the slab go into the matrix, the right-hand side is the
 * gathered one (get_full_array_in_place, :230). Returns 0, -1 on the sanity check of :176-177. */
int orc_damping_add(orc_csr *matrix, double *b_RHS, double alpha, double problem_weight, double norm_power,
                    int32_t compression_type, int32_t nx, int32_t ny, int32_t nz,
                    int32_t nsmaller, int32_t nelements, const double *column_weight, const double *model,
                    const double *model_ref, int32_t param_shift, int32_t wavelet_domain,
                    const double *local_weight, double *cost) {
  const int32_t ntot = nx * ny * nz;
  double *model_diff = (double *)malloc(sizeof(double) * (size_t)ntot);
  for (int32_t i = 0; i < ntot; ++i) {                                         /* :117-126 */
    double dm = model[i] - model_ref[i];
    model_diff[i] = (column_weight[i] != 0.0) ? dm / column_weight[i] : 0.0;
  }
  if (compression_type > 0 && wavelet_domain)                                  /* :128-142 */
    orc_forward_wavelet(model_diff, nx, ny, nz, compression_type);
  const int32_t row_beg = matrix->nl_current_all + 1;                          /* :145 */
  orc_csr_add_empty_rows(matrix, nsmaller);                                    /* :151 */
  for (int32_t i = 0; i < nelements; ++i) {                                    /* :154-168 */
    const int32_t p = nsmaller + i;
    double value = alpha * problem_weight;
    if (norm_power != 2.0)
      value = value * ((model_diff[p] != 0.0) ? pow(fabs(model_diff[p]), norm_power / 2.0 - 1.0) : 1.0);
    if (local_weight) value = value * local_weight[p];
    if (orc_csr_add(matrix, value, param_shift + i + 1) != 0) { free(model_diff); return -2; }
    if (orc_csr_new_row(matrix) != 0) { free(model_diff); return -2; }
  }
  orc_csr_add_empty_rows(matrix, ntot - nelements - nsmaller);                 /* :171 */
  const int32_t row_end = matrix->nl_current_all;
  if (row_end - row_beg + 1 != ntot) { free(model_diff); return -1; }          /* :176-177 */
  double c = 0.0;
  for (int32_t i = 0; i < ntot; ++i) {                                         /* :213-230 (all ranks' slabs) */
    double b = -alpha * problem_weight * model_diff[i];
    if (norm_power != 2.0)
      b = b * ((model_diff[i] != 0.0) ? pow(fabs(model_diff[i]), norm_power / 2.0 - 1.0) : 1.0);
    if (local_weight) b = b * local_weight[i];
    b_RHS[row_beg - 1 + i] = b;
    c += b * b;                                                                /* :190 */
  }
  if (cost) *cost = c;
  free(model_diff);
  return 0;
}

/* grad_get_par, src/inversion/gradient.F90:175-225: zero outside the domain. */
static double orc_grad_par(const double *val, int32_t nx, int32_t ny, int32_t nz, int32_t i, int32_t j, int32_t k) {
  if (i == nx + 1 || j == ny + 1 || k == nz + 1) return 0.0;
  if (i == 0 || j == 0 || k == 0) return 0.0;
  return val[(i - 1) + (size_t)(j - 1) * nx + (size_t)(k - 1) * nx * ny];
}
/* grad_grid_get_ind, src/inversion/grid.F90:409-426: 1-based, -1 outside. */
static int32_t orc_get_ind(int32_t nx, int32_t ny, int32_t nz, int32_t i, int32_t j, int32_t k) {
  if (i < 1 || i > nx || j < 1 || j > ny || k < 1 || k > nz) return -1;
  return i + (j - 1) * nx + (k - 1) * nx * ny;
}
/* get_grad, src/inversion/gradient.F90:71-170: type 0 backward, 1 forward, 2 central. */
static void orc_get_grad(const double *val, int32_t nx, int32_t ny, int32_t nz, const double *dX, const double *dY,
                         const double *dZ, int32_t i, int32_t j, int32_t k, int type, double g[3]) {
#define PAR(a, b, c) orc_grad_par(val, nx, ny, nz, (a), (b), (c))
  if (type == 0) {
    g[0] = (PAR(i, j, k) - PAR(i - 1, j, k)) / dX[i - 1];
    g[1] = (PAR(i, j, k) - PAR(i, j - 1, k)) / dY[j - 1];
    g[2] = (PAR(i, j, k) - PAR(i, j, k - 1)) / dZ[k - 1];
  } else if (type == 1) {
    g[0] = (PAR(i + 1, j, k) - PAR(i, j, k)) / dX[i - 1];
    g[1] = (PAR(i, j + 1, k) - PAR(i, j, k)) / dY[j - 1];
    g[2] = (PAR(i, j, k + 1) - PAR(i, j, k)) / dZ[k - 1];
  } else {
    g[0] = (PAR(i + 1, j, k) - PAR(i - 1, j, k)) / 2.0 / dX[i - 1];
    g[1] = (PAR(i, j + 1, k) - PAR(i, j - 1, k)) / 2.0 / dY[j - 1];
    g[2] = (PAR(i, j, k + 1) - PAR(i, j, k - 1)) / 2.0 / dZ[k - 1];
  }
#undef PAR
}

/* damping_gradient_add, src/inversion/damping_gradient.F90:93-203. val_full: the full model component;
 * column_weight: full-grid array, the slab's entries are used (column_weight(ind - nsmaller) there);
 * local_weight(nx*ny*nz). Returns 0, -1 wrong direction. */
int orc_damping_gradient_add(orc_csr *matrix, double *b_RHS, double beta, double problem_weight,
                             int32_t nx, int32_t ny, int32_t nz, const double *dX, const double *dY, const double *dZ,
                             int32_t nsmaller, int32_t nelements, const double *val_full, const double *column_weight,
                             const double *local_weight, int32_t param_shift, int32_t direction, double *cost) {
  if (direction < 1 || direction > 3) return -1;
  double c = 0.0;
  int32_t p = 0;
  for (int32_t k = 1; k <= nz; ++k)
    for (int32_t j = 1; j <= ny; ++j)
      for (int32_t i = 1; i <= nx; ++i) {
        p++;
        double g[3], delta, gradient_val;
        int32_t ind[2];
        orc_get_grad(val_full, nx, ny, nz, dX, dY, dZ, i, j, k, 1, g);
        if (direction == 1) {
          delta = dX[i - 1];
          if (i == nx) { orc_csr_new_row(matrix); continue; }
          ind[0] = orc_get_ind(nx, ny, nz, i + 1, j, k);
        } else if (direction == 2) {
          delta = dY[j - 1];
          if (j == ny) { orc_csr_new_row(matrix); continue; }
          ind[0] = orc_get_ind(nx, ny, nz, i, j + 1, k);
        } else {
          delta = dZ[k - 1];
          if (k == nz) { orc_csr_new_row(matrix); continue; }
          ind[0] = orc_get_ind(nx, ny, nz, i, j, k + 1);
        }
        ind[1] = orc_get_ind(nx, ny, nz, i, j, k);
        gradient_val = g[direction - 1];
        double val[2];
        val[0] = 1.0 / delta;
        val[1] = -val[0];
        for (int l = 0; l < 2; ++l) {                                          /* :177-184 */
          if (ind[l] > nsmaller && ind[l] <= nsmaller + nelements) {
            const int32_t loc = ind[l] - nsmaller;
            const double v = val[l] * problem_weight * beta * column_weight[ind[l] - 1] * local_weight[p - 1];
            if (orc_csr_add(matrix, v, param_shift + loc) != 0) return -2;
          }
        }
        if (orc_csr_new_row(matrix) != 0) return -2;
        b_RHS[matrix->nl_current_all - 1] = -problem_weight * beta * gradient_val * local_weight[p - 1];   /* :189 */
        c = c + gradient_val * gradient_val;                                   /* :192 */
      }
  if (cost) *cost = c;
  return 0;
}

/* cross_gradient_calculate with add = .true., vec_field_type = 0 (src/inversion/cross_gradient.F90:220-391),
 * calculate_tau (:455-567) and calculate_tau_backward (:676-740). der_type 1 (forward) or 2 (central).
 * model1/model2: full models; column_weight1/2: full-grid arrays (the slab's entries are used);
 * cost[3]; cross_grad(nx*ny*nz) may be NULL. Returns 0, -1 unsupported derivative type. */
typedef struct { double val[3]; double dm1[4][3]; double dm2[4][3]; int32_t ind[4][3]; } orc_tau;

static void orc_tau_zero(orc_tau *t) { memset(t, 0, sizeof(*t)); }

static void orc_calc_tau(const double *m1, const double *m2, int32_t nx, int32_t ny, int32_t nz, const double *dX,
                         const double *dY, const double *dZ, int32_t i, int32_t j, int32_t k, int der_type, orc_tau *t) {
  double g1[3], g2[3];
  orc_tau_zero(t);
  orc_get_grad(m1, nx, ny, nz, dX, dY, dZ, i, j, k, der_type == 1 ? 1 : 2, g1);   /* get_der_type: 1 FWD, 2 CNT */
  orc_get_grad(m2, nx, ny, nz, dX, dY, dZ, i, j, k, der_type == 1 ? 1 : 2, g2);
  t->val[0] = g1[1] * g2[2] - g1[2] * g2[1];                                   /* vector.f90 cross_product */
  t->val[1] = g1[2] * g2[0] - g1[0] * g2[2];
  t->val[2] = g1[0] * g2[1] - g1[1] * g2[0];
  double sx = dX[i - 1], sy = dY[j - 1], sz = dZ[k - 1];
  if (der_type != 1) { sx = 2.0 * sx; sy = 2.0 * sy; sz = 2.0 * sz; }
#define IND(a, b, c) orc_get_ind(nx, ny, nz, (a), (b), (c))
  /* x */
  t->dm1[0][0] = g2[2] / sy;  t->dm2[0][0] = -g1[2] / sy;
  t->dm1[1][0] = -g2[1] / sz; t->dm2[1][0] = g1[1] / sz;
  t->ind[0][0] = IND(i, j + 1, k); t->ind[1][0] = IND(i, j, k + 1);
  if (der_type == 1) {
    t->dm1[2][0] = -(g2[2] / sy - g2[1] / sz); t->dm2[2][0] = -(g1[1] / sz - g1[2] / sy);
    t->ind[2][0] = IND(i, j, k);
  } else {
    t->dm1[2][0] = -t->dm1[0][0]; t->dm2[2][0] = -t->dm2[0][0];
    t->dm1[3][0] = -t->dm1[1][0]; t->dm2[3][0] = -t->dm2[1][0];
    t->ind[2][0] = IND(i, j - 1, k); t->ind[3][0] = IND(i, j, k - 1);
  }
  /* y */
  t->dm1[0][1] = -g2[2] / sx; t->dm2[0][1] = g1[2] / sx;
  t->dm1[1][1] = g2[0] / sz;  t->dm2[1][1] = -g1[0] / sz;
  t->ind[0][1] = IND(i + 1, j, k); t->ind[1][1] = IND(i, j, k + 1);
  if (der_type == 1) {
    t->dm1[2][1] = -(g2[0] / sz - g2[2] / sx); t->dm2[2][1] = -(g1[2] / sx - g1[0] / sz);
    t->ind[2][1] = IND(i, j, k);
  } else {
    t->dm1[2][1] = -t->dm1[0][1]; t->dm2[2][1] = -t->dm2[0][1];
    t->dm1[3][1] = -t->dm1[1][1]; t->dm2[3][1] = -t->dm2[1][1];
    t->ind[2][1] = IND(i - 1, j, k); t->ind[3][1] = IND(i, j, k - 1);
  }
  /* z */
  t->dm1[0][2] = g2[1] / sx;  t->dm2[0][2] = -g1[1] / sx;
  t->dm1[1][2] = -g2[0] / sy; t->dm2[1][2] = g1[0] / sy;
  t->ind[0][2] = IND(i + 1, j, k); t->ind[1][2] = IND(i, j + 1, k);
  if (der_type == 1) {
    t->dm1[2][2] = -(g2[1] / sx - g2[0] / sy); t->dm2[2][2] = -(g1[0] / sy - g1[1] / sx);
    t->ind[2][2] = IND(i, j, k);
  } else {
    t->dm1[2][2] = -t->dm1[0][2]; t->dm2[2][2] = -t->dm2[0][2];
    t->dm1[3][2] = -t->dm1[1][2]; t->dm2[3][2] = -t->dm2[1][2];
    t->ind[2][2] = IND(i - 1, j, k); t->ind[3][2] = IND(i, j - 1, k);
  }
}

static void orc_calc_tau_backward(const double *m1, const double *m2, int32_t nx, int32_t ny, int32_t nz,
                                  const double *dX, const double *dY, const double *dZ, int32_t i, int32_t j, int32_t k,
                                  orc_tau *t) {
  double g1[3], g2[3];
  orc_tau_zero(t);
  orc_get_grad(m1, nx, ny, nz, dX, dY, dZ, i, j, k, 0, g1);
  orc_get_grad(m2, nx, ny, nz, dX, dY, dZ, i, j, k, 0, g2);
  t->val[0] = g1[1] * g2[2] - g1[2] * g2[1];
  t->val[1] = g1[2] * g2[0] - g1[0] * g2[2];
  t->val[2] = g1[0] * g2[1] - g1[1] * g2[0];
  const double sx = dX[i - 1], sy = dY[j - 1], sz = dZ[k - 1];
  t->dm1[0][0] = -g2[2] / sy; t->dm1[1][0] = g2[1] / sz;  t->dm1[2][0] = g2[2] / sy - g2[1] / sz;
  t->dm2[0][0] = g1[2] / sy;  t->dm2[1][0] = -g1[1] / sz; t->dm2[2][0] = g1[1] / sz - g1[2] / sy;
  t->ind[0][0] = IND(i, j - 1, k); t->ind[1][0] = IND(i, j, k - 1); t->ind[2][0] = IND(i, j, k);
  t->dm1[0][1] = g2[2] / sx;  t->dm1[1][1] = -g2[0] / sz; t->dm1[2][1] = g2[0] / sz - g2[2] / sx;
  t->dm2[0][1] = -g1[2] / sx; t->dm2[1][1] = g1[0] / sz;  t->dm2[2][1] = g1[2] / sx - g1[0] / sz;
  t->ind[0][1] = IND(i - 1, j, k); t->ind[1][1] = IND(i, j, k - 1); t->ind[2][1] = IND(i, j, k);
  t->dm1[0][2] = -g2[1] / sx; t->dm1[1][2] = g2[0] / sy;  t->dm1[2][2] = g2[1] / sx - g2[0] / sy;
  t->dm2[0][2] = g1[1] / sx;  t->dm2[1][2] = -g1[0] / sy; t->dm2[2][2] = g1[0] / sy - g1[1] / sx;
  t->ind[0][2] = IND(i - 1, j, k); t->ind[1][2] = IND(i, j - 1, k); t->ind[2][2] = IND(i, j, k);
#undef IND
}

int orc_cross_gradient_calculate(orc_csr *matrix, double *b_RHS, int32_t nx, int32_t ny, int32_t nz,
                                 const double *dX, const double *dY, const double *dZ,
                                 int32_t nsmaller, int32_t nparams_loc, const double *model1, const double *model2,
                                 const double *column_weight1, const double *column_weight2,
                                 int32_t der_type, double glob_weight, const int32_t keep_model_constant[2],
                                 double cost[3], double *cross_grad, int64_t *nnz_out, int32_t *nl_nonempty_out) {
  if (der_type != 1 && der_type != 2) return -1;                               /* :281-283 */
  const int nderiv = (der_type == 1) ? 3 : 4;                                  /* :203-215 */
  int64_t nnz = 0;
  int32_t nl_nonempty = 0, p = 0;
  cost[0] = cost[1] = cost[2] = 0.0;
  for (int32_t k = 1; k <= nz; ++k)
    for (int32_t j = 1; j <= ny; ++j)
      for (int32_t i = 1; i <= nx; ++i) {
        p++;
        orc_tau t;
        const int left = (i == 1 || j == 1 || k == 1), right = (i == nx || j == ny || k == nz);
        if (left && right) orc_tau_zero(&t);                                   /* :262-266 */
        else if (der_type == 2 && left) orc_calc_tau(model1, model2, nx, ny, nz, dX, dY, dZ, i, j, k, 1, &t);
        else if (right) orc_calc_tau_backward(model1, model2, nx, ny, nz, dX, dY, dZ, i, j, k, &t);
        else orc_calc_tau(model1, model2, nx, ny, nz, dX, dY, dZ, i, j, k, der_type, &t);
        if (keep_model_constant[0]) memset(t.dm1, 0, sizeof(t.dm1));           /* :286-287 */
        if (keep_model_constant[1]) memset(t.dm2, 0, sizeof(t.dm2));
        if (cross_grad) cross_grad[p - 1] = sqrt(t.val[0] * t.val[0] + t.val[1] * t.val[1] + t.val[2] * t.val[2]);
        for (int c = 0; c < 3; ++c) cost[c] = cost[c] + t.val[c] * t.val[c];   /* :298-300 */
        for (int c = 0; c < 3; ++c) {                                          /* :305-371 */
          int included = 0;
          for (int l = 0; l < nderiv; ++l) {
            int32_t ind = t.ind[l][c];
            if (ind > nsmaller && ind <= nsmaller + nparams_loc) {
              const double val1 = t.dm1[l][c] * column_weight1[ind - 1] * glob_weight;
              const double val2 = t.dm2[l][c] * column_weight2[ind - 1] * glob_weight;
              ind = ind - nsmaller;
              if (orc_csr_add(matrix, val1, ind) != 0) return -2;
              if (orc_csr_add(matrix, val2, ind + nparams_loc) != 0) return -2;
              nnz += 2;
              included = 1;
            }
          }
          if (orc_csr_new_row(matrix) != 0) return -2;
          b_RHS[matrix->nl_current_all - 1] = -t.val[c] * glob_weight;
          if (included) nl_nonempty++;
        }
      }
  if (keep_model_constant[0] && keep_model_constant[1]) nnz = 0;               /* :378-383 */
  else if (keep_model_constant[0] || keep_model_constant[1]) nnz = nnz / 2;
  if (nnz_out) *nnz_out = nnz;
  if (nl_nonempty_out) *nl_nonempty_out = nl_nonempty;
  return 0;
}

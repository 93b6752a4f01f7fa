/*
 * tfx_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the Tomofast-x inversion hot path, used only as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under tomofast-x_b200/ may link or call it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout, e.g. src/inversion/sparse_matrix.f90).
 *
 * Parity pinning: the reference cannot be built in this environment (no
 * Fortran compiler, no MPI), so the oracle is pinned by the reference's own
 * unit-test known answers (tests/test_oracle_*.py):
 *   - Haar3D identity nnz == 46656          (tests_wavelet_compression.f90:179)
 *   - norm preservation, exact inverse      (tests_wavelet_compression.f90:187-326)
 *   - wavelet-domain mat-vec equality       (tests_wavelet_compression.f90:70-135)
 *   - six LSQR solutions                    (tests_lsqr.f90:71-624)
 *   - CSR build / mult_vector / col norms   (tests_sparse_matrix.f90:39-113)
 * The forward kernels (graviprism_z, magprism/sharmbox), the compression
 * pipeline, part_mult_vector, lsqr_solve_sensit and ADMM are NOT covered by any
 * reference test: for those the oracle is "parity unpinned" (line-by-line
 * restatement only).
 */
#ifndef TFX_ORACLE_H
#define TFX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- t_sparse_matrix (src/inversion/sparse_matrix.f90:31-98) ------------- */
typedef struct orc_csr {
  int64_t nnz;                 /* predicted number of non-zeros              */
  int64_t nel;                 /* actual number of stored elements           */
  int64_t nel_last;
  int32_t nl;                  /* total number of rows                       */
  int32_t nl_nonempty;
  int32_t nl_nonempty_allocated;
  int32_t nl_current;          /* excludes empty rows                        */
  int32_t nl_current_all;
  int32_t ncolumns;
  float   *sa;                 /* values, real(4)                            */
  int32_t *ija;                /* 1-based column indices                     */
  int64_t *ijl;                /* 1-based row starts                         */
  int32_t *rowptr;             /* stored row -> global row (1-based)         */
  int32_t finalized;
} orc_csr;

orc_csr *orc_csr_new(int32_t nl, int32_t ncolumns, int64_t nnz, int32_t nl_empty);
void     orc_csr_free(orc_csr *m);
void     orc_csr_reset(orc_csr *m);
int      orc_csr_add(orc_csr *m, double value, int32_t column);
int      orc_csr_add_row(orc_csr *m, int32_t nel_add, const float *values, const int32_t *columns);
int      orc_csr_new_row(orc_csr *m);
void     orc_csr_add_empty_rows(orc_csr *m, int32_t nrows);
int      orc_csr_finalize(orc_csr *m);
void     orc_csr_mult_vector(const orc_csr *m, const double *x, double *b);
void     orc_csr_add_mult_vector(const orc_csr *m, const double *x, double *b);
int      orc_csr_part_mult_vector(const orc_csr *m, int32_t nelements, const double *x, int32_t ndata,
                                  double *b, int32_t line_start, int32_t param_shift);
void     orc_csr_trans_mult_vector(const orc_csr *m, const double *x, double *b);
void     orc_csr_add_trans_mult_vector(const orc_csr *m, const double *x, double *b);
void     orc_csr_normalize_columns(orc_csr *m, double *column_norm);

/* ---- wavelet_transform (src/utils/wavelet_transform.F90) ----------------- */
void orc_haar3d(double *s, int n1, int n2, int n3);
void orc_ihaar3d(double *s, int n1, int n2, int n3);
void orc_daubd43d(double *s, int n1, int n2, int n3);
void orc_idaubd43d(double *s, int n1, int n2, int n3);
int  orc_forward_wavelet(double *s, int n1, int n2, int n3, int wavelet_type);
int  orc_inverse_wavelet(double *s, int n1, int n2, int n3, int wavelet_type);

/* ---- lsqr_solver (src/inversion/lsqr_solver2.F90) ------------------------ */
/* r_hist (may be NULL) receives r = phibar/b1 after each executed iteration.
 * Returns 0 ok, <0 on the reference's fatal conditions. *iters = iter-1.     */
int orc_lsqr_solve(int32_t nlines, int32_t nelements, int32_t niter, double rmin, double gamma,
                   const orc_csr *matrix, double *u, double *x, double *r_hist, int32_t *iters);
int orc_lsqr_solve_sensit(int32_t nlines, int32_t ncolumns, int32_t niter, double rmin, double gamma,
                          double target_misfit, const orc_csr *matrix_sensit, const orc_csr *matrix_cons,
                          double *u, double *x, const int32_t solve_problem[2], int32_t nelements,
                          int32_t nx, int32_t ny, int32_t nz, int32_t ncomponents,
                          int32_t compression_type, int32_t wavelet_domain,
                          double *r_hist, int32_t *iters);

/* ---- forward kernels ----------------------------------------------------- */
/* src/forward/gravmag/grav/gravity_field.f90:131-195. Returns 0, or 1/2 on the
 * "data coincides with grid boundary" aborts.                                */
int  orc_graviprism_z(int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                      const double *Z1, const double *Z2, double xd, double yd, double zd, double *lineZ);
void orc_gradiprism_zz(int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                       const double *Z1, const double *Z2, double xd, double yd, double zd, double *lineZZ);
int  orc_gradiprism_full(int32_t n, const double *X1, const double *X2, const double *Y1, const double *Y2,
                         const double *Z1, const double *Z2, double xd, double yd, double zd, double *lines);
/* src/forward/gravmag/mag/magnetic_field.f90 */
void orc_dircos(double incl, double decl, double azim, double *a, double *b, double *c);
int  orc_sharmbox(double x0, double y0, double z0, double x1, double y1, double z1,
                  double x2, double y2, double z2, double tsx[3], double tsy[3], double tsz[3]);
int  orc_magprism(int32_t n, int32_t nmodel_comp, int32_t ndata_comp,
                  const double *X1, const double *X2, const double *Y1, const double *Y2,
                  const double *Z1, const double *Z2, double xd, double yd, double zd,
                  double mi, double md, double theta, double intensity, double *sensit_line);

/* ---- row compression (src/forward/gravmag/sensitivity_gravmag.F90:222-311)  */
/* line: one (k,d) sensitivity line, ALREADY multiplied by the column weight.
 * On return line holds the wavelet-transformed row (if compression_type>0).
 * cols are 1-based. Returns nel.                                             */
int32_t orc_compress_row(double *line, int32_t nx, int32_t ny, int32_t nz, int32_t compression_type,
                         int32_t nel_compressed, int32_t *cols, float *vals,
                         double *threshold_out, double *cost_full_out, double *cost_discarded_out);

/* ---- depth weight (src/forward/gravmag/weights_gravmag.f90:46-250) -------- */
int orc_depth_weight(int32_t type, int32_t n, const double *X1, const double *X2, const double *Y1,
                     const double *Y2, const double *Z1, const double *Z2, int32_t ndata,
                     const double *xd, const double *yd, const double *zd,
                     double power, double beta, double Z0, double *column_weight);

/* ---- ADMM (src/inversion/admm_method.F90:70-134) -------------------------- */
void orc_admm_iterate(int32_t n, int32_t nlithos, const double *xmin, const double *xmax,
                      const double *x, double *z, double *u, double *x0);

/* ---- constraint-matrix producers ("parity unpinned": no reference test covers them) ----
 * damping.F90:97-261, damping_gradient.F90:93-203 (+ gradient.F90:71-225, grid.F90:409-426),
 * cross_gradient.F90:220-391,455-567,676-740. Full-grid input arrays; the rank's column slab is
 * (nsmaller, nelements). b_RHS = the constraint part of the right-hand side. */
int orc_damping_add(orc_csr *matrix, double *b_RHS, double alpha, double problem_weight, double norm_power,
                    int32_t compression_type, int32_t nx, int32_t ny, int32_t nz,
                    int32_t nsmaller, int32_t nelements, const double *column_weight, const double *model,
                    const double *model_ref, int32_t param_shift, int32_t wavelet_domain,
                    const double *local_weight, double *cost);
int orc_damping_gradient_add(orc_csr *matrix, double *b_RHS, double beta, double problem_weight,
                             int32_t nx, int32_t ny, int32_t nz, const double *dX, const double *dY, const double *dZ,
                             int32_t nsmaller, int32_t nelements, const double *val_full, const double *column_weight,
                             const double *local_weight, int32_t param_shift, int32_t direction, double *cost);
int orc_cross_gradient_calculate(orc_csr *matrix, double *b_RHS, int32_t nx, int32_t ny, int32_t nz,
                                 const double *dX, const double *dY, const double *dZ,
                                 int32_t nsmaller, int32_t nparams_loc, const double *model1, const double *model2,
                                 const double *column_weight1, const double *column_weight2,
                                 int32_t der_type, double glob_weight, const int32_t keep_model_constant[2],
                                 double cost[3], double *cross_grad, int64_t *nnz_out, int32_t *nl_nonempty_out);

double orc_norm2(int64_t n, const double *x);

#ifdef __cplusplus
}
#endif
#endif

/* Minimal C host of libtfx: the calling sequence a Tomofast-x style driver uses for one solve -- build the
 * sensitivity rows and a damping block with the t_sparse_matrix builder calls, finalize (device mirror), solve,
 * read the residual history. Mirrors the reference's unit test of the solver (src/tests/tests_lsqr.f90:71-118).
 *
 *   gcc -std=c99 -Iinclude examples/host_lsqr.c -Ltomofast-x_b200 -ltfx -L/usr/local/cuda/lib64 -lcudart \
 *       -Wl,-rpath,$PWD/tomofast-x_b200 -o host_lsqr && ./host_lsqr          (needs a GPU to run)
 */
#include <stdio.h>
#include <stdlib.h>

#include "tfx.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    int rc_ = (call);                                                      \
    if (rc_ != 0) {                                                        \
      fprintf(stderr, "libtfx: %s (code %d)\n", tfx_last_error(), rc_);    \
      return 1; /* the Fortran shim calls exit_MPI here */                 \
    }                                                                      \
  } while (0)

int main(void) {
  enum { NROWS = 3, NCOLS = 3 };
  /* A = [1 0 0; 0 4 0; 0 0 9], b = (1, 4, 9)  ->  x = (1, 1, 1) */
  const double diag[NROWS] = {1.0, 4.0, 9.0};
  double u[NROWS] = {1.0, 4.0, 9.0}, x[NCOLS] = {0.0, 0.0, 0.0}, hist[64];
  tfx_matrix *A = NULL;
  int32_t iters = 0, fused = 0;
  int i;

  CHECK(tfx_init(-1));
  CHECK(tfx_sparse_matrix_initialize(&A, NROWS, NCOLS, NROWS, 0, 0));
  for (i = 0; i < NROWS; ++i) {
    CHECK(tfx_sparse_matrix_add(A, diag[i], i + 1, 0)); /* 1-based column, like the reference */
    CHECK(tfx_sparse_matrix_new_row(A, 0));
  }
  CHECK(tfx_sparse_matrix_finalize(A, 0));
  CHECK(tfx_lsqr_solve(NROWS, NCOLS, 50, 1.0e-13, 0.0, A, u, x, 0)); /* u is overwritten, x receives the solution */
  CHECK(tfx_lsqr_last_history(hist, 64, &iters, &fused));
  printf("x = %.12f %.12f %.12f after %d iterations, r = %.3e\n", x[0], x[1], x[2], (int)iters,
         iters > 0 ? hist[iters - 1] : 0.0);
  CHECK(tfx_sparse_matrix_destroy(A));
  CHECK(tfx_finalize());
  return 0;
}

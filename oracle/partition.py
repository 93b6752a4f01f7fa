"""CPU ORACLE (test infrastructure, NOT product code): the reference's re-partitioning of the
sensitivity kernel from data-sharded rows to column slabs, restated with plain Python / numpy loops.

    get_load_balancing_nelements   src/forward/gravmag/sensitivity_gravmag.F90:470-524
    read_sensitivity_kernel        src/forward/gravmag/sensitivity_gravmag.F90:648-883 (partition + index shift)
    calculate_nelements_at_cpu     src/utils/parallel_tools.f90:46-63

Parity unpinned: no reference test covers these routines; this is a line-by-line restatement.
"""
import numpy as np


def calculate_nelements_at_cpu(nelements_total, myrank, nbproc):
    n = nelements_total // nbproc
    if myrank + 1 <= nelements_total - n * nbproc:
        n += 1
    return n


def get_load_balancing_nelements(sensit_nnz, nbproc):
    """Returns (nnz_at_cpu_new, nelements_at_cpu_new); raises like the reference's exit_MPI calls."""
    sensit_nnz = [int(v) for v in sensit_nnz]
    nelements_total = len(sensit_nnz)
    nnz_total = sum(sensit_nnz)                                    # :485-489
    best = [nnz_total // nbproc] * nbproc                          # :491
    best[nbproc - 1] += nnz_total % nbproc                         # :493
    cpu = 1
    nnz_new = 0
    nelements_new = 0
    nnz_at_cpu_new = [0] * nbproc
    nelements_at_cpu_new = [0] * nbproc
    sum_sensit_nnz = 0
    for p in range(1, nelements_total + 1):                        # :501-514
        nnz_new += sensit_nnz[p - 1]
        sum_sensit_nnz += sensit_nnz[p - 1]
        nelements_new += 1
        if (cpu <= nbproc and sum_sensit_nnz >= sum(best[:cpu]) and cpu < nbproc) or p == nelements_total:
            if cpu > nbproc:
                raise RuntimeError("Wrong cpu in get_load_balancing_nelements!")
            nnz_at_cpu_new[cpu - 1] = nnz_new
            nelements_at_cpu_new[cpu - 1] = nelements_new
            nnz_new = 0
            nelements_new = 0
            cpu += 1
    if cpu != nbproc + 1:                                          # :517-519
        raise RuntimeError("Wrong cpu in get_load_balancing_nelements!")
    if sum(nnz_at_cpu_new) != nnz_total:                           # :521-523
        raise RuntimeError("Wrong nnz_at_cpu_new in get_load_balancing_nelements!")
    return np.array(nnz_at_cpu_new, dtype=np.int64), np.array(nelements_at_cpu_new, dtype=np.int32)


def column_slab(rows, N, nmc, nelements_at_cpu, myrank, problem_slot):
    """rows: {global_row(1-based): (cols, vals)} of the kernel with columns (k-1)*N + p (1-based, no problem shift,
    the content of the stream files). Returns the same dict for rank `myrank`'s slab with the LOCAL column
    index of read_sensitivity_kernel (:759-846): cells nsmaller < p <= nsmaller + nelements keep
    p + param_shift(slot) + (k-1)*nelements - nsmaller, param_shift = (0, nelements*nmc) (:685-686)."""
    cum = np.concatenate([[0], np.cumsum(nelements_at_cpu)])
    nsmaller, nel = int(cum[myrank]), int(nelements_at_cpu[myrank])
    param_shift = (problem_slot - 1) * nel * nmc
    out = {}
    for r, (cols, vals) in rows.items():
        cols = np.asarray(cols, dtype=np.int64)
        k = (cols - 1) // N                                         # 0-based model component
        p = (cols - 1) % N + 1                                      # 1-based cell
        sel = (p > nsmaller) & (p <= nsmaller + nel)
        if sel.any():
            out[r] = ((p[sel] + param_shift + k[sel] * nel - nsmaller).astype(np.int32), np.asarray(vals)[sel])
    return out

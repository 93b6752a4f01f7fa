"""CPU ORACLE (test infrastructure, NOT product code): the reference's sensitivity file formats restated in
numpy. Big-endian like the reference build (-fconvert=big-endian, Makefile:51).

    stream file   src/forward/gravmag/sensitivity_gravmag.F90:143-148 (name), :183 (header), :306-309 (records)
    _meta.txt     :359-376      _nnz  :381-392      _weight  :415-465
    reader        :648-883 (read_sensitivity_kernel), :974-1037 (metadata)

Parity unpinned: the reference holds no fixture of these files; this follows the cited write/read statements.
"""
import os

import numpy as np

SUFFIX = {1: "grav", 2: "magn"}                                    # :57


def rank_file(path, problem_type, nbproc, rank):
    return os.path.join(path, "sensit_%s_%d_%d" % (SUFFIX[problem_type], nbproc, rank))


def write_rank_file(path, problem_type, nbproc, rank, ndata, N, records):
    """records: list of (idata, k, d, cols(1-based cells), vals(f32)) in the order (idata, d, k)."""
    ndata_loc = len(set(r[0] for r in records))
    with open(rank_file(path, problem_type, nbproc, rank), "wb") as f:
        f.write(np.array([ndata_loc, ndata, N, rank, nbproc], dtype=">i4").tobytes())
        for idata, k, d, cols, vals in records:
            f.write(np.array([idata, len(cols), k, d], dtype=">i4").tobytes())
            if len(cols):
                f.write(np.asarray(cols, dtype=">i4").tobytes())
                f.write(np.asarray(vals, dtype=">f4").tobytes())


def read_rank_file(path, problem_type, nbproc, rank):
    """-> (header tuple, records) with records as in write_rank_file."""
    raw = open(rank_file(path, problem_type, nbproc, rank), "rb").read()
    hdr = np.frombuffer(raw, dtype=">i4", count=5)
    off = 20
    recs = []
    while off < len(raw):
        idata, nel, k, d = (int(v) for v in np.frombuffer(raw, dtype=">i4", count=4, offset=off))
        off += 16
        cols = np.frombuffer(raw, dtype=">i4", count=nel, offset=off).astype(np.int32); off += 4 * nel
        vals = np.frombuffer(raw, dtype=">f4", count=nel, offset=off).astype(np.float32); off += 4 * nel
        recs.append((idata, k, d, cols, vals))
    return tuple(int(v) for v in hdr), recs


def write_meta(path, problem_type, nx, ny, nz, ndata, nbproc, weight_type, compression_type, comp_error, nmc, ndc,
               nnz_total, gfortran_style=False):
    with open(os.path.join(path, "sensit_%s_meta.txt" % SUFFIX[problem_type]), "w") as f:
        if gfortran_style:      # what list-directed output of gfortran looks like (wide fields)
            f.write("%12d%12d%12d%12d\n" % (nx, ny, nz, ndata))
            f.write("%12d%12d%12d\n" % (nbproc, 4, weight_type))
            f.write("%12d   %.16E     \n" % (compression_type, comp_error))
            f.write("%12d%12d\n" % (nmc, ndc))
            f.write("%21d\n" % nnz_total)
        else:
            f.write("%d %d %d %d\n%d 4 %d\n%d %r\n%d %d\n%d\n" % (nx, ny, nz, ndata, nbproc, weight_type,
                                                                  compression_type, float(comp_error), nmc, ndc, nnz_total))


def write_nnz(path, problem_type, sensit_nnz):
    with open(os.path.join(path, "sensit_%s_nnz" % SUFFIX[problem_type]), "wb") as f:
        f.write(np.array([len(sensit_nnz)], dtype=">i4").tobytes())
        f.write(np.asarray(sensit_nnz, dtype=">i4").tobytes())


def read_nnz(path, problem_type):
    raw = open(os.path.join(path, "sensit_%s_nnz" % SUFFIX[problem_type]), "rb").read()
    n = int(np.frombuffer(raw, dtype=">i4", count=1)[0])
    return np.frombuffer(raw, dtype=">i4", count=n, offset=4).astype(np.int32)


def write_weight(path, problem_type, cw):
    with open(os.path.join(path, "sensit_%s_weight" % SUFFIX[problem_type]), "wb") as f:
        f.write(np.array([len(cw)], dtype=">i4").tobytes())
        f.write(np.asarray(cw, dtype=">f8").tobytes())


def read_weight(path, problem_type):
    raw = open(os.path.join(path, "sensit_%s_weight" % SUFFIX[problem_type]), "rb").read()
    n = int(np.frombuffer(raw, dtype=">i4", count=1)[0])
    return np.frombuffer(raw, dtype=">f8", count=n, offset=4).astype(np.float64)

"""Device sensitivity assembly (csrc/assembly.cu, csrc/sensit.cu) vs the oracle.

The forward kernels are sums of 8 signed corner terms with heavy cancellation; CUDA's log/atan2 differ
from glibc's in the last ulp, which the cancellation amplifies. Tolerances are therefore relative to the
largest |value| of the line (stated per test). No reference test covers these kernels ("parity unpinned"):
the oracle is a line-by-line restatement."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.synth import make_problem

pytestmark = pytest.mark.gpu


def test_gravity_lines_vs_oracle(oracle):
    pb = make_problem(nx=12, ny=10, nz=6, ndata=9)
    got = tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)
    for i in range(pb.ndata):
        want = oracle.graviprism_z(pb.grid, *(float(a[i]) for a in pb.data_xyz))
        err = np.abs(got[i, 0, 0] - want).max() / np.abs(want).max()
        assert err < 1e-12, err


def test_gravity_shared_nodes_is_bit_identical_to_per_cell_evaluation(oracle):
    """Structured grids evaluate every prism corner term once per grid NODE and share it between the (up to 8) cells
    around it (csrc/assembly.cu grav_lines_nodes_kernel); the per-cell sums keep the reference's order, so the lines are
    bit-identical to the per-cell kernel. Grid sizes that are not multiples of the 32 x 8 x 8 tile."""
    pb = make_problem(nx=37, ny=11, nz=9, ndata=5)
    try:
        tfx.set_option("grav_shared_nodes", 0)
        per_cell = tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)
    finally:
        tfx.set_option("grav_shared_nodes", 1)
    shared = tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)
    assert np.array_equal(shared, per_cell)
    want = oracle.graviprism_z(pb.grid, *(float(a[2]) for a in pb.data_xyz))
    assert np.abs(shared[2, 0, 0] - want).max() / np.abs(want).max() < 1e-12


@pytest.mark.parametrize("nmc,ndc", [(1, 1), (3, 1), (1, 3), (3, 3)])
def test_magnetic_shared_terms_are_bit_identical_to_per_cell_evaluation(oracle, nmc, ndc):
    """Structured grids evaluate sharmbox's corner terms (atan2) once per node and its edge terms (log of a ratio) once
    per edge (csrc/assembly.cu mag_lines_nodes_kernel); orders and signs of the per-cell sums are the reference's.
    One station sits inside a cell (six-sub-prism branch, magnetic_field.f90:139-224)."""
    pb = make_problem(nx=35, ny=10, nz=9, ndata=4, problem_type=2, nmodel_components=nmc, ndata_components=ndc)
    x, y, z = (a.copy() for a in pb.data_xyz)
    x[1], y[1], z[1] = 1237.3, 451.9, 161.0                      # inside cell (12, 4, 3)
    try:
        tfx.set_option("mag_shared_nodes", 0)
        per_cell = tfx.sensit_lines(pb.par, pb.grid, (x, y, z))
    finally:
        tfx.set_option("mag_shared_nodes", 1)
    shared = tfx.sensit_lines(pb.par, pb.grid, (x, y, z))
    assert np.array_equal(shared, per_cell)
    p = pb.par
    want = oracle.magprism(pb.grid, float(x[1]), float(y[1]), float(z[1]), nmc, ndc, p.mi, p.md, p.theta, p.intensity)
    assert np.abs(shared[1] - want).max() / np.abs(want).max() < 1e-11


def test_gravity_lines_on_an_unstructured_set_of_boxes(oracle):
    """Arbitrary per-cell boxes (the reference stores X1..Z2 per cell, gravity_field.f90:151-156): a grid whose boxes
    do not share their faces falls back to the per-cell kernel."""
    pb = make_problem(nx=6, ny=5, nz=4, ndata=4)
    rng = np.random.default_rng(9)
    grid = [a.copy() for a in pb.grid]
    grid[1] -= rng.uniform(0.5, 3.0, grid[1].size)          # X2: gaps between neighbours
    grid[5] += rng.uniform(0.0, 4.0, grid[5].size)          # Z2: overlapping piles
    got = tfx.sensit_lines(pb.par, grid, pb.data_xyz)
    for i in range(pb.ndata):
        want = oracle.graviprism_z(grid, *(float(a[i]) for a in pb.data_xyz))
        assert np.abs(got[i, 0, 0] - want).max() / np.abs(want).max() < 1e-12


@pytest.mark.parametrize("nmc,ndc", [(1, 1), (3, 1), (1, 3), (3, 3)])
def test_magnetic_lines_vs_oracle(oracle, nmc, ndc):
    pb = make_problem(nx=9, ny=8, nz=5, ndata=6, problem_type=2, nmodel_components=nmc, ndata_components=ndc)
    got = tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)
    p = pb.par
    for i in range(pb.ndata):
        want = oracle.magprism(pb.grid, *(float(a[i]) for a in pb.data_xyz), nmc, ndc, p.mi, p.md, p.theta, p.intensity)
        err = np.abs(got[i] - want).max() / np.abs(want).max()
        assert err < 1e-11, err


def test_magnetic_station_inside_cell(oracle):
    # borehole branch: six sub-prisms around a void (magnetic_field.f90:139-224)
    pb = make_problem(nx=5, ny=5, nz=4, ndata=1, problem_type=2, nmodel_components=3, ndata_components=1)
    pb.data_xyz = (np.array([237.3]), np.array([151.9]), np.array([61.0]))
    got = tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)
    p = pb.par
    want = oracle.magprism(pb.grid, 237.3, 151.9, 61.0, 3, 1, p.mi, p.md, p.theta, p.intensity)
    assert np.abs(got[0] - want).max() / np.abs(want).max() < 1e-11


def test_station_on_cell_boundary_aborts():
    pb = make_problem(nx=4, ny=4, nz=2, ndata=1)
    # station on the line through a cell edge along x (y and z on cell faces): for the cells ahead
    # Rs + XX = |XX| + XX = 0 (gravity_field.f90:173-178)
    pb.data_xyz = (np.array([150.0]), np.array([100.0]), np.array([50.0]))
    with pytest.raises(tfx.TfxError, match="coincides with model grid boundary"):
        tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)


def _rows(sa, ija, ijl, rowptr):
    return {int(rowptr[i]): (ija[ijl[i] - 1:ijl[i + 1] - 1], sa[ijl[i] - 1:ijl[i + 1] - 1]) for i in range(len(rowptr))}


def test_dense_assembly_vs_oracle(oracle):
    pb = make_problem(nx=10, ny=9, nz=5, ndata=20, compression_type=0)
    S, nnz_col, cerr, tot = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    assert S.storage_kind() == 1 and tot == pb.ndata * pb.N and np.all(nnz_col == pb.ndata)
    So = pb.oracle_matrix(oracle)
    want = _rows(*So.arrays())
    got = _rows(*S.export())
    for r in want:
        assert np.array_equal(got[r][0], want[r][0])
        # f32 values: equal up to f32 rounding of a value that differs by ~1e-13 relative in f64
        assert np.allclose(got[r][1], want[r][1], rtol=3e-7, atol=1e-7 * np.abs(want[r][1]).max())


@pytest.mark.parametrize("ctype", [1, 2])
def test_compressed_assembly_vs_oracle(oracle, ctype):
    pb = make_problem(nx=12, ny=10, nz=6, ndata=12, compression_type=ctype, rate=0.2)
    S, nnz_col, cerr, tot = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    assert S.storage_kind() == 0
    So = pb.oracle_matrix(oracle)
    want = _rows(*So.arrays())
    got = _rows(*S.export())
    assert set(want) == set(got)
    mismatched = 0
    for r in want:
        if np.array_equal(got[r][0], want[r][0]):
            assert np.allclose(got[r][1], want[r][1], rtol=3e-6, atol=1e-6 * np.abs(want[r][1]).max())
        else:
            # a coefficient within rounding of the threshold may flip: at most a couple per row
            mismatched += len(set(got[r][0]) ^ set(want[r][0]))
    assert mismatched <= 2 * pb.ndata, mismatched
    assert abs(tot - So.nel) <= mismatched
    # the products agree to compression-noise-free precision when the patterns agree
    x = pb.model_scaled_w(oracle)
    d_want = So.mult_vector(x)
    d_got = S.mult_vector(x)
    assert np.allclose(d_got, d_want, rtol=1e-6, atol=1e-8 * np.abs(d_want).max())


@pytest.mark.parametrize("ctype", [1, 2])
def test_compressed_assembly_many_chunks_vs_oracle(oracle, ctype):
    """A grid that spans several 4096-element compaction chunks and several CTAs per line in the select passes
    (csrc/sensit.cu k_sel_* / k_cmp_*): pattern, values and per-cell counts vs the oracle."""
    pb = make_problem(nx=48, ny=40, nz=10, ndata=6, compression_type=ctype, rate=0.07)
    S, nnz_col, cerr, tot = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    So = pb.oracle_matrix(oracle)
    want = _rows(*So.arrays())
    got = _rows(*S.export())
    assert set(want) == set(got)
    mismatched = 0
    for r in want:
        assert np.all(np.diff(got[r][0]) > 0)                 # columns ascending inside a row
        if np.array_equal(got[r][0], want[r][0]):
            assert np.allclose(got[r][1], want[r][1], rtol=3e-6, atol=1e-6 * np.abs(want[r][1]).max())
        else:
            mismatched += len(set(got[r][0]) ^ set(want[r][0]))
    assert mismatched <= 2 * pb.ndata, mismatched
    assert abs(tot - So.nel) <= mismatched
    assert int(nnz_col.sum()) == tot
    # compression error = mean over the lines of sqrt(discarded cost / full cost) (sensitivity_gravmag.F90:283-285, :346-353)
    errs = []
    for i in range(pb.ndata):
        line = oracle.graviprism_z(pb.grid, *(float(a[i]) for a in pb.data_xyz)) * pb.cw
        r = oracle.compress_row(line, pb.par.nx, pb.par.ny, pb.par.nz, ctype, pb.nel_compressed)
        errs.append(np.sqrt(r["cost_discarded"] / r["cost_full"]))
    assert cerr == pytest.approx(np.mean(errs), rel=1e-3)


def test_kth_select_fallback_passes_give_the_same_rows():
    """The k-th order statistic resolves its low 39 bits on a candidate list; a bucket that does not fit the list goes on
    with radix passes over the whole line. Forcing that path (sensit_cand_cap = 1) must not change a single entry."""
    pb = make_problem(nx=40, ny=32, nz=16, ndata=5, compression_type=1, rate=0.05)
    rows_a, nnz_a, cerr_a, tot_a = tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    S_a, _, _, _ = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    try:
        tfx.set_option("sensit_cand_cap", 1)
        S_b, nnz_b, cerr_b, tot_b = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    finally:
        tfx.set_option("sensit_cand_cap", 0)
    assert tot_a == tot_b and cerr_a == cerr_b and np.array_equal(nnz_a, nnz_b)
    for x, y in zip(S_a.export(), S_b.export()):
        assert np.array_equal(x, y)
    # exactly nel_compressed entries per row unless values tie at the threshold
    nel = int(pb.par.compression_rate * pb.N)
    assert tot_a <= nel * pb.ndata and tot_a >= (nel - 2) * pb.ndata


def test_pinned_grid_gives_the_same_rows():
    """tfx_grid_pin keeps the device copy of the grid for calls that pass the same host arrays; other arrays (even with
    equal content) are uploaded as before."""
    pb = make_problem(nx=20, ny=12, nz=8, ndata=7, compression_type=1, rate=0.1)
    S0, nnz0, cerr0, tot0 = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    try:
        pinned = tfx.grid_pin(pb.grid)
        S1, nnz1, cerr1, tot1 = tfx.calculate_sensit(pb.par, pinned, pb.data_xyz, pb.cw, pb.dw)
        rows, nnz2, cerr2, tot2 = tfx.sensit_assemble_rows(pb.par, pinned, pb.data_xyz, pb.cw, pb.dw)
        other = tuple(np.array(a, copy=True) for a in pb.grid)
        S3, nnz3, cerr3, tot3 = tfx.calculate_sensit(pb.par, other, pb.data_xyz, pb.cw, pb.dw)
    finally:
        tfx.grid_unpin()
    assert tot0 == tot1 == tot2 == tot3 and cerr0 == cerr1 == cerr2 == cerr3
    assert np.array_equal(nnz0, nnz1) and np.array_equal(nnz0, nnz2) and np.array_equal(nnz0, nnz3)
    for x, y, z in zip(S0.export(), S1.export(), S3.export()):
        assert np.array_equal(x, y) and np.array_equal(x, z)


def test_magnetic_compressed_assembly_three_components(oracle):
    # config D shape in miniature: magnetisation model (3 comps), TMI data, columns shifted to problem 2
    pb = make_problem(nx=8, ny=7, nz=4, ndata=6, compression_type=2, rate=0.3, problem_type=2, nmodel_components=3)
    S, nnz_col, cerr, tot = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    So = pb.oracle_matrix(oracle)
    want = _rows(*So.arrays())
    got = _rows(*S.export())
    bad = sum(len(set(got[r][0]) ^ set(want[r][0])) for r in want)
    assert bad <= 2 * pb.ndata * 3
    assert int(min(want[1][0])) > 3 * pb.N          # param_shift(2) = nelements * ncomponents
    x = pb.model_scaled_w(oracle)
    assert np.allclose(S.mult_vector(x), So.mult_vector(x), rtol=1e-6, atol=1e-9)


# ---- rows sharded by data -> nnz-balanced column slabs (csrc/sensit_dist.cu) ---------------------------------
def _full_rows_no_shift(pb, oracle):
    """The content of the reference's stream files: rows with columns (k-1)*N + p, weights applied."""
    shift = pb.par.param_shift
    pb.par.param_shift = 0
    try:
        So = pb.oracle_matrix(oracle)
    finally:
        pb.par.param_shift = shift
    return _rows(*So.arrays())


@pytest.mark.parametrize("case", ["grav_haar", "mag3_d4"])
@pytest.mark.parametrize("nbproc", [1, 2, 3])
def test_repartition_single_process_vs_oracle(oracle, case, nbproc):
    from oracle import partition as orp
    if case == "grav_haar":
        pb = make_problem(nx=12, ny=10, nz=6, ndata=11, compression_type=1, rate=0.2)
        slot = 1
    else:
        pb = make_problem(nx=8, ny=7, nz=4, ndata=6, compression_type=2, rate=0.3, problem_type=2, nmodel_components=3)
        slot = 2
    nmc = pb.par.nmodel_components
    want_full = _full_rows_no_shift(pb, oracle)
    rows, nnz_col, cerr, tot = tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    assert rows.info() == (0, pb.ndata, tot)
    # same counts as the one-call assembly
    S1, nnz_col1, cerr1, tot1 = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    assert tot == tot1 and np.array_equal(nnz_col, nnz_col1) and cerr == cerr1
    nnz_at, nel_at = tfx.get_load_balancing_nelements(nnz_col, nbproc)
    assert np.array_equal(nel_at, orp.get_load_balancing_nelements(nnz_col, nbproc)[1])
    full_got = _rows(*S1.export())
    total = 0
    for r in range(nbproc):
        if r > 0:
            rows, _, _, _ = tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
        M = tfx.sensit_repartition(rows, slot, nel_at, r, nbproc)
        assert M.get_ncolumns() == 2 * nmc * nel_at[r] and M.get_total_row_number() == pb.ndata
        got = _rows(*M.export())
        # exact against the device's own full matrix (same kernels, same values), restricted by the oracle's rule
        shift = pb.par.param_shift
        dev_rows = {k: (c - shift, v) for k, (c, v) in full_got.items()}
        want_dev = orp.column_slab(dev_rows, pb.N, nmc, nel_at, r, slot)
        assert set(got) == set(want_dev)
        for k in got:
            assert np.array_equal(got[k][0], want_dev[k][0]) and np.array_equal(got[k][1], want_dev[k][1])
        # and against the oracle's matrix up to threshold flips (as in test_compressed_assembly_vs_oracle)
        want = orp.column_slab(want_full, pb.N, nmc, nel_at, r, slot)
        bad = sum(len(set(got.get(k, ([], []))[0]) ^ set(want.get(k, ([], []))[0])) for k in set(got) | set(want))
        assert bad <= 2 * pb.ndata * nmc
        total += M.get_number_elements()
        assert abs(M.get_number_elements() - nnz_at[r]) == 0
    assert total == tot


def test_repartition_even_split_and_lsqr_slab(oracle):
    """Even column split (parallel_tools.f90:46-63) and a product on the slab against the oracle's."""
    from oracle import partition as orp
    pb = make_problem(nx=12, ny=10, nz=6, ndata=11, compression_type=1, rate=0.25)
    nbproc = 4
    nel_at = np.array([orp.calculate_nelements_at_cpu(pb.N, r, nbproc) for r in range(nbproc)], dtype=np.int32)
    So = pb.oracle_matrix(oracle)
    x = pb.model_scaled_w(oracle)
    d_want = So.mult_vector(x)
    d_got = np.zeros(pb.ndata)
    cum = np.concatenate([[0], np.cumsum(nel_at)])
    for r in range(nbproc):
        rows, _, _, _ = tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
        M = tfx.sensit_repartition(rows, 1, nel_at, r, nbproc)
        xl = np.zeros(2 * nel_at[r])
        xl[:nel_at[r]] = x[cum[r]:cum[r + 1]]
        d_got += M.mult_vector(xl)                 # the MPI_Allreduce of model.F90:293, summed here
    assert np.allclose(d_got, d_want, rtol=1e-6, atol=1e-8 * np.abs(d_want).max())


def test_repartition_argument_checks():
    pb = make_problem(nx=6, ny=5, nz=4, ndata=5, compression_type=1, rate=0.3)
    rows, _, _, _ = tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    with pytest.raises(tfx.TfxError, match="does not sum"):
        tfx.sensit_repartition(rows, 1, np.array([10, 10], dtype=np.int32), 0, 2)
    with pytest.raises(tfx.TfxError, match="problem_slot"):
        tfx.sensit_repartition(rows, 3, np.array([pb.N], dtype=np.int32), 0, 1)
    with pytest.raises(tfx.TfxError, match="communicator"):
        tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw, myrank=0, nbproc=2)


def test_rows_apply_weights_equals_weighting_at_assembly():
    """Unit-weight rows (the file content) + apply_weights == rows assembled with the weights: both are
    real(line, 4) * real(pw * dw, 4) in real(4) (sensitivity_gravmag.F90:265, :837-843)."""
    pb = make_problem(nx=9, ny=8, nz=5, ndata=8, compression_type=1, rate=0.3)
    dw = np.linspace(0.3, 1.7, pb.ndata).reshape(pb.ndata, 1)
    pb.par.problem_weight = 0.625
    rows_w, _, _, _ = tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, dw)
    A = tfx.sensit_repartition(rows_w, 1, [pb.N])
    pb.par.problem_weight = 1.0
    rows_u, _, _, _ = tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, np.ones_like(dw))
    tfx.sensit_rows_apply_weights(rows_u, 0.625, dw)
    with pytest.raises(tfx.TfxError, match="already carry weights"):
        tfx.sensit_rows_apply_weights(rows_u, 0.625, dw)
    B = tfx.sensit_repartition(rows_u, 1, [pb.N])
    a, b = A.export(), B.export()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)

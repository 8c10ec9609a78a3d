"""Trajectory -> transition counts.

CPU part: the numpy oracle (oracle/traj_oracle.py) against golden vectors produced by the unmodified reference
functions transition_matrix_add1 / count_2D (tests/golden/make_traj_golden.py).
GPU part (-m gpu): mcd_traj_counts / mcd_traj_counts_rad through mcdiff_b200.extract against the goldens and the
oracle -- integer work, every comparison is exact."""
import os

import numpy as np
import pytest

from oracle import traj_oracle as TO
from tests.conftest import GOLDEN


@pytest.fixture(scope="module")
def gold_traj():
    return np.load(os.path.join(GOLDEN, "traj_golden.npz"))


# ---------------------------------------------------------------- CPU: oracle pinned against the reference
def test_oracle_matches_reference_counts(gold_traj):
    G = gold_traj
    got = TO.transition_counts(G["pbc_z"], G["pbc_edges"], list(G["pbc_shifts"]), period=float(G["pbc_zpbc"]))
    assert np.array_equal(got, G["pbc_counts"])
    assert got.sum(axis=(1, 2)).tolist() == [(400 - s) * 7 for s in G["pbc_shifts"]]
    got = TO.transition_counts(G["nu_x"], G["nu_edges"], list(G["nu_shifts"]))
    assert np.array_equal(got, G["nu_counts"])
    for l, s in enumerate(G["nu_shifts"]):                     # the single-trajectory form with the reference's name
        A = TO.transition_matrix_add1(np.zeros((10, 10), int), G["nu_x"], G["nu_edges"], shift=int(s))
        assert np.array_equal(A, G["nu_counts"][l])


def test_oracle_matches_reference_radial_cube(gold_traj):
    G = gold_traj
    for l, s in enumerate(G["rad_shifts"]):
        B = np.zeros(G["rad_counts"].shape[1:], int)
        TO.count_2D(B, G["rad_X"], G["rad_Y"], G["rad_Z"], G["rad_edges"], G["rad_redges"], shift=int(s))
        assert np.array_equal(B, G["rad_counts"][l])
    assert G["rad_counts"][:, -1].sum() > 0                     # displacements beyond the last radial edge occur


# ---------------------------------------------------------------- GPU: the device path, bit-exact
@pytest.mark.gpu
def test_device_counts_equal_reference(engine, gold_traj):
    from mcdiff_b200 import extract as X
    G = gold_traj
    got = X.transition_counts(G["pbc_z"], G["pbc_edges"], G["pbc_shifts"], period=float(G["pbc_zpbc"]), engine=engine)
    assert got.dtype == np.int64 and np.array_equal(got, G["pbc_counts"])
    got = X.transition_counts(G["nu_x"], G["nu_edges"], G["nu_shifts"], engine=engine)
    assert np.array_equal(got, G["nu_counts"])
    # the reference's own call shape: accumulate into an existing matrix, one shift
    A = G["nu_counts"][0].copy()
    A = X.transition_matrix_add1(A, G["nu_x"], G["nu_edges"], shift=int(G["nu_shifts"][0]), engine=engine)
    assert np.array_equal(A, 2 * G["nu_counts"][0])
    with pytest.raises(AssertionError):                        # lib/tools/extract.py:73
        X.transition_matrix_add1(np.zeros((10, 10), int), G["nu_x"], G["nu_edges"], shift=300, engine=engine)


@pytest.mark.gpu
def test_device_radial_cube_equals_reference(engine, gold_traj):
    from mcdiff_b200 import extract as X
    G = gold_traj
    got = X.count_2D_device(G["rad_X"], G["rad_Y"], G["rad_Z"], G["rad_edges"], G["rad_redges"], G["rad_shifts"],
                            engine=engine).cpu().numpy()
    assert np.array_equal(got, G["rad_counts"])
    B = np.zeros(G["rad_counts"].shape[1:], int)
    for i in range(G["rad_X"].shape[1]):                       # per particle, as the reference is called
        B = X.count_2D(B, G["rad_X"][:, i], G["rad_Y"][:, i], G["rad_Z"][:, i], G["rad_edges"], G["rad_redges"],
                       shift=int(G["rad_shifts"][1]), engine=engine)
    assert np.array_equal(B, G["rad_counts"][1])


@pytest.mark.gpu
@pytest.mark.parametrize("T,M,nedges,nshift", [(5000, 1, 101, 4), (3001, 37, 101, 11), (257, 130, 9, 3),
                                               (4000, 16, 201, 2), (2000, 5, 300, 3), (1500, 3, 1000, 2)])
def test_device_counts_equal_oracle_shapes(engine, T, M, nedges, nshift):
    """Ragged sizes (element counts that are not multiples of 4, odd particle counts -> unaligned partner words),
    one-byte and two-byte bin indices (255 / 256+ edges), shared-memory and L2 histograms (237+ edges), several shift
    groups (11 shifts at 100 bins), NaN / infinite / outside coordinates, non-uniform edges."""
    from mcdiff_b200 import extract as X
    rng = np.random.default_rng(T + M)
    edges = np.sort(rng.uniform(-30, 30, size=nedges)) if nedges in (9, 300) else np.linspace(-30, 30, nedges)
    z = np.cumsum(rng.normal(0, 0.9, size=(T, M)), axis=0) + rng.uniform(-25, 25, size=M)
    z[rng.integers(0, T, 5), rng.integers(0, M, 5)] = np.nan
    z[rng.integers(0, T, 3), rng.integers(0, M, 3)] = np.inf
    z[rng.integers(0, T, 3), rng.integers(0, M, 3)] = -np.inf
    z[rng.integers(0, T, 20), rng.integers(0, M, 20)] = edges[rng.integers(0, nedges, 20)]    # exactly on an edge
    shifts = sorted(set([1, T - 1] + list(rng.integers(1, T // 2, size=nshift))))
    for period in (None, 47.3):
        want = TO.transition_counts(z, edges, shifts, period=period)
        got = X.transition_counts(z if M > 1 else z[:, 0], edges, shifts, period=period, engine=engine)
        assert np.array_equal(got, want), (T, M, nedges, period)
        assert got.sum(axis=(1, 2)).tolist() == [(T - s) * M for s in shifts]


@pytest.mark.gpu
def test_device_counts_large_properties_and_files(engine, tmp_path):
    """2^24 coordinates (larger than the oracle loop is pleasant for): every pair is counted exactly once, the counts
    of a time-reversed trajectory are the transposed counts, an unaligned device view gives the same matrix; and the
    example driver (extract_Tmat) writes files the Transitions reader accepts."""
    import torch
    from mcdiff_b200 import extract as X
    from mcdiff_b200.transitions import Transitions
    T, M, nbins, zpbc = 1 << 16, 256, 100, 50.0
    g = torch.Generator(device="cuda").manual_seed(5)
    z = torch.cumsum(torch.randn((T, M), generator=g, device="cuda", dtype=torch.float64) * 0.4, dim=0)
    edges = np.arange(-nbins * (zpbc / nbins) / 2, (nbins + 1) * (zpbc / nbins) / 2., zpbc / nbins)
    shifts = [10, 20, 30, 40]
    c = X.transition_counts_device(z, edges, shifts, period=zpbc, engine=engine)
    assert c.sum(dim=(1, 2)).tolist() == [(T - s) * M for s in shifts]
    assert int(c[:, 0].sum()) == 0 and int(c[:, -1].sum()) == 0            # folded coordinates never leave the box
    cr = X.transition_counts_device(torch.flip(z, dims=[0]), edges, shifts, period=zpbc, engine=engine)
    assert torch.equal(cr, c.transpose(1, 2))
    zz = torch.empty(T * M + 1, device="cuda", dtype=torch.float64)
    zz[1:] = z.reshape(-1)                                                 # 8-byte aligned only: scalar digitize path
    cu = X.transition_counts_device(zz[1:].view(T, M), edges, shifts, period=zpbc, engine=engine)
    assert torch.equal(cu, c)
    sub = z[:4096, :8].contiguous()
    want = TO.transition_counts(sub.cpu().numpy(), edges, shifts, period=zpbc)
    assert np.array_equal(X.transition_counts(sub, edges, shifts, period=zpbc, engine=engine), want)
    files = X.extract_Tmat(sub, zpbc, nbins, shifts, str(tmp_path), dt=1.0, engine=engine)
    data = Transitions(files)
    assert data.dim_trans == nbins and list(data.list_lt) == [10.0, 20.0, 30.0, 40.0]
    assert np.array_equal(np.asarray(data.list_trans).astype(np.int64), want[:, 1:-1, 1:-1])


@pytest.mark.gpu
def test_device_counts_argument_errors(engine):
    import ctypes as C
    import torch
    from mcdiff_b200 import _capi as K
    z = torch.zeros(100, device="cuda", dtype=torch.float64)
    e = torch.linspace(-1, 1, 5, device="cuda", dtype=torch.float64)
    out = torch.zeros((1, 6, 6), device="cuda", dtype=torch.int64)

    def call(nframes, shift, nsh=1):
        sh = (C.c_int32 * 1)(shift)
        return engine.lib.mcd_traj_counts(engine.handle, C.c_void_p(z.data_ptr()), nframes, 1, 0.0,
                                          C.c_void_p(e.data_ptr()), 5, sh, nsh, C.c_void_p(out.data_ptr()), None)
    assert call(100, 100) == K.load().mcd_traj_counts(engine.handle, None, 100, 1, 0.0, None, 5, None, 1, None, None) < 0
    assert call(100, 0) < 0 and call(1, 1) < 0 and call(100, 1, 65) < 0 and call(100, 99) == 0
    torch.cuda.synchronize()
    assert int(out.sum()) == 1


# ---------------------------------------------------------------- CPU: host-side argument handling (no device call is reached)
def test_host_argument_checks():
    from mcdiff_b200 import extract as X
    with pytest.raises(ValueError):
        X._check_edges([0.0, 2.0, 1.0])                         # np.digitize's other numbering is never used by the reference
    with pytest.raises(ValueError):
        X._check_edges([[0.0, 1.0]])
    assert X._check_edges(np.arange(5)).dtype == np.float64
    with pytest.raises(AssertionError):
        X._shifts([5], 5)                                       # lib/tools/extract.py:73: shift < len(x)
    with pytest.raises(AssertionError):
        X._shifts([0], 5)
    assert X._shifts(3, 10).tolist() == [3]
    A = np.zeros((4, 4), int)
    with pytest.raises(AssertionError):
        X.transition_matrix_add1(A, np.zeros(10), np.arange(5.0), shift=1)      # len(A) == len(edges) + 1
    with pytest.raises(AssertionError):
        X.count_2D(np.zeros((3, 4, 4)), None, None, None, np.arange(3.0), np.arange(4.0))   # B.shape[0] == len(redges)

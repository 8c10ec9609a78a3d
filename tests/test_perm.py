"""Propagators with absorbing boundaries for the permeability analyses.

CPU part: the oracle (oracle/perm_oracle.py) against golden matrices made with the unmodified reference's
init_rate_matrix / construct_rate_matrix_from_F_D (tests/golden/make_perm_golden.py); the host decomposition of a cut
matrix into (v, w, sink); the chain body compiled for the host (tests/emul) with absorbing rates against the goldens.
GPU part (-m gpu): mcd_propagator_batch through mcdiff_b200.permeability against the same goldens.
Tolerance (north_star): 1e-9 relative on every propagator entry >= 1e-10, 1e-14 absolute below."""
import os

import numpy as np
import pytest

from oracle import mcdiff_oracle as O
from oracle import perm_oracle as PO
from tests.conftest import GOLDEN

SIDES = {0: "both", 1: "left", 2: "right", 3: None}


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "perm_golden.npz"))


def cuts(G):
    for k in range(int(G["ncuts"])):
        st, end, side = (int(a) for a in G["cut%d_args" % k])
        yield k, st, end, SIDES[side]


def assert_prop(got, want):
    big = want >= 1e-10
    np.testing.assert_allclose(got[big], want[big], rtol=1e-9, atol=0)
    np.testing.assert_allclose(got[~big], want[~big], rtol=0, atol=1e-14)


# ---------------------------------------------------------------- CPU
def test_oracle_rate_matrices_and_numbers_match_reference(gold):
    G = gold
    F, D, dx, dt = G["F"], G["D"], float(G["dx"]), float(G["dt"])
    for k, st, end, side in cuts(G):
        R = PO.construct_rate_matrix_from_F_D(F - F.min(), D, dx, dt, pbc=False, st=st, end=end, side=side)
        assert np.array_equal(R, G["cut%d_rate" % k])
    rho, cdf = PO.distribution_mfpt(F, D, dx, dt, int(G["mfpt_b"][0]), int(G["mfpt_b"][1]), G["mfpt_trange"])
    np.testing.assert_allclose(rho, G["mfpt_rho"], rtol=1e-13)
    np.testing.assert_allclose(cdf, G["mfpt_cdf"], rtol=1e-13, atol=1e-16)
    st, end = (int(a) for a in G["part_st_end"])
    for tag, pbc in (("pbc", True), ("ref", False)):
        data = PO.partitioning_transient(F, D, pbc, dt, st, end, G["part_times"], G["part_p0"])
        np.testing.assert_allclose(data, G["part_%s_data" % tag], rtol=1e-12, atol=1e-18)


def test_cut_matrix_is_reflecting_generator_minus_sink(gold):
    """The host decomposition the device call relies on: cut matrix == nopbc generator of the kept bins - diag(sink)."""
    from mcdiff_b200.permeability import absorbing_cut, profiles_from_F_D
    G = gold
    F, D, dx, dt = G["F"], G["D"], float(G["dx"]), float(G["dt"])
    v, w = profiles_from_F_D(F - F.min(), D, dx, dt)
    for k, st, end, side in cuts(G):
        vs, ws, sink = absorbing_cut(v, w, st, end, side)
        R = O.rate_matrix(vs, ws, False) - np.diag(sink)
        np.testing.assert_allclose(R, G["cut%d_rate" % k], rtol=1e-15, atol=0)
        assert (sink[1:-1] == 0).all() and (sink[0] > 0) == (side in ("left", "both"))


def test_chain_body_with_absorbing_rates_on_the_host(gold):
    """tests/emul: the CUDA chain body (tensor-core squaring route, host execution policy) with sink != 0."""
    from mcdiff_b200.permeability import absorbing_cut, profiles_from_F_D
    from tests.emul import harness as H
    G = gold
    F, D, dx, dt = G["F"], G["D"], float(G["dx"]), float(G["dt"])
    v, w = profiles_from_F_D(F - F.min(), D, dx, dt)
    for k, st, end, side in list(cuts(G))[:2]:
        vs, ws, sink = absorbing_cut(v, w, st, end, side)
        m = len(vs)
        prob = H.HostProblem(np.zeros((1, m, m), dtype=np.int64), np.array([1.0]), False)
        times = G["cut_times"]
        P, status = H.propagator_batch(prob, np.tile(vs, (3, 1)), np.tile(ws, (3, 1)), np.tile(sink, (3, 1)), times)
        assert not status.any()
        assert_prop(P, G["cut%d_prop" % k])


# ---------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_device_propagators_equal_reference(engine, gold):
    from mcdiff_b200 import permeability as PM
    G = gold
    F, D, dx, dt = G["F"], G["D"], float(G["dx"]), float(G["dt"])
    for k, st, end, side in cuts(G):
        P = PM.propagators(F - F.min(), D, dx, dt, G["cut_times"], pbc=False, st=st, end=end, side=side, engine=engine)
        assert_prop(P, G["cut%d_prop" % k])
    # the two uncut forms: periodic and reflecting (lib/utils.py:214-217)
    for pbc in (True, False):
        P = PM.propagators(F, D, dx, dt, [7.5, 90.0], pbc=pbc, engine=engine)
        R = PO.construct_rate_matrix_from_F_D(F, D, dx, dt, pbc=pbc)
        for i, t in enumerate((7.5, 90.0)):
            assert_prop(P[i], PO.expm(R * t))
        np.testing.assert_allclose(P.sum(axis=1), 1.0, rtol=1e-12)          # no absorption: columns stay normalised
    with pytest.raises(NotImplementedError):
        PM.propagators(F, D, dx, dt, [1.0], pbc=True, st=3, end=20, side="both", engine=engine)


@pytest.mark.gpu
def test_device_exit_time_distribution_and_transient(engine, gold, tmp_path):
    from mcdiff_b200 import permeability as PM
    G = gold
    F, D, dx, dt = G["F"], G["D"], float(G["dx"]), float(G["dt"])
    name = str(tmp_path / "t_rho.dat")
    rho, cdf = PM.distribution_mfpt(F, D, dx, dt, int(G["mfpt_b"][0]), int(G["mfpt_b"][1]), G["mfpt_trange"],
                                    filename=name, engine=engine)
    # rho picks the two corner entries P[0, k], P[-1, k]: below 1e-10 only the absolute tolerance of a propagator entry
    # applies (scipy's own value at t = 1 is -2e-18)
    np.testing.assert_allclose(rho, G["mfpt_rho"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(cdf, G["mfpt_cdf"], rtol=1e-9, atol=1e-13)    # 1 - sum P cancels at early times
    assert (np.diff(cdf[2:]) > 0).all() and cdf[-1] <= 1 and (rho > -1e-13).all()
    np.testing.assert_allclose(np.loadtxt(name), np.column_stack([G["mfpt_trange"], rho, cdf]), rtol=1e-15)
    st, end = (int(a) for a in G["part_st_end"])
    for tag, pbc in (("pbc", True), ("ref", False)):
        data, prob = PM.partitioning_transient(F, D, pbc, dt, st, end, G["part_times"], G["part_p0"], engine=engine)
        np.testing.assert_allclose(prob, G["part_%s_prob" % tag], rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(data, G["part_%s_data" % tag], rtol=1e-8, atol=1e-13)   # fluxes are differences


@pytest.mark.gpu
@pytest.mark.parametrize("n", [24, 100, 150])
def test_device_propagator_batch_mixed_items(engine, n):
    """Every item of a batch has its own profile, absorbing rates and time (small-profile kernel is not involved: the
    likelihood kernels serve all sizes; n = 150 runs on the blocked engine), against scipy's expm entry by entry."""
    from mcdiff_b200.engine import PropagatorProblem
    rng = np.random.default_rng(n)
    B = 9
    v = rng.normal(0, 0.8, size=(B, n))
    w = rng.normal(-0.5, 0.3, size=(B, n - 1))
    sink = np.zeros((B, n))
    sink[:, 0] = rng.uniform(0, 1.5, size=B)
    sink[:, -1] = rng.uniform(0, 1.5, size=B)
    sink[3] = rng.uniform(0, 0.2, size=n)                                   # a leak in every bin
    sink[4] = 0.0
    times = rng.uniform(0.2, 60.0, size=B)
    prob = PropagatorProblem(engine, n, False)
    P, status = prob.propagators(v, w, times, sink=sink)
    assert not status.any()
    for b in range(B):
        want = PO.expm(times[b] * (O.rate_matrix(v[b], w[b], False) - np.diag(sink[b])))
        assert_prop(P[b], want)
    vbad = v.copy()
    vbad[2, 5] = np.nan
    _, status = prob.propagators(vbad, w, times, sink=sink)
    assert status[2] != 0 and not np.delete(status, 2).any()


def test_host_cut_argument_checks():
    from mcdiff_b200.permeability import absorbing_cut
    v, w = np.zeros(10), np.zeros(10)
    for st, end in ((-2, 5), (3, 11), (4, 5), (9, 10)):          # the asserts of lib/utils.py:66-68
        with pytest.raises(AssertionError):
            absorbing_cut(v, w, st, end, "both")
    with pytest.raises(AssertionError):
        absorbing_cut(v, np.zeros(9), 2, 8, "both")              # the cut is taken from the periodic matrix
    vs, ws, sink = absorbing_cut(v, w, None, None, None)         # st = -1, end = n, both ends reflective
    assert len(vs) == 10 and len(ws) == 9 and not sink.any()

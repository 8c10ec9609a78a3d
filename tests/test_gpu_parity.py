"""GPU parity tests: the CUDA path (through the C ABI) against the reference's golden
fixtures and the CPU oracle.

Tolerances (north star): log-likelihood and propagator within 1e-9 relative of the
reference's numpy/scipy path in fp64; accept/reject decisions bit-exact for a host-supplied
tape of proposals and uniforms.
Propagator entries are compared norm-wise (|dP| <= 1e-9 max|P|) and entry-wise above 1e-6:
expm itself is only accurate to ~1e-16 absolute, so tiny entries carry no 1e-9 relative
information in the reference either.
"""
import numpy as np
import pytest

from tests.helpers import PULL_TAGS, REPLAY_TAGS, replay_case

pytestmark = pytest.mark.gpu

RTOL = 1e-9


METHODS = {"eigen": 1, "squaring": 2, "squaring_dfma": 3}


def _params(case, nsteps, use_tape=1, seed=0, refresh=64, chain_offset=0, method=0):
    from mcdiff_b200 import _capi as K
    mp = K.McParams()
    mp.num_mc_update = case["nmc_update"]
    mp.dtemp = case["dtemp"]
    mp.min_lt = float(np.min(case["lagtimes"]))
    mp.move_timezero = int(case["move_timezero"])
    mp.sample_every = 100
    mp.refresh_every = refresh
    mp.use_tape = use_tape
    mp.seed = seed
    mp.chain_offset = chain_offset
    mp.tape_stride = 0
    mp.method = method
    return mp


def _problem(engine, case):
    from mcdiff_b200.engine import DeviceProblem
    return DeviceProblem(engine, case["counts"], case["lagtimes"], case["pbc"], v_basis=case["v_basis"],
                         w_basis=case["w_basis"], spring_k=case["k"], string_vecs=case["svecs"])


@pytest.mark.parametrize("method", ["eigen", "squaring", "squaring_dfma"])
def test_known_answers_K1_K4(engine, gold_ll, method):
    """The four archived `initial log-likelihood` values of the reference's example logs."""
    from mcdiff_b200.engine import DeviceProblem
    G = gold_ll
    m = METHODS[method]
    pA = DeviceProblem(engine, G["A_counts"], G["A_lags"], True)
    pM = DeviceProblem(engine, G["M_counts"], G["M_lags"], True)
    assert pA.squaring_ok and pA.in_smem
    llA, _, stA = pA.loglike(G["K_v"][:2], G["K_w"][:2], method=m)
    llM, _, stM = pM.loglike(G["K_v"][2:], G["K_w"][2:], method=m)
    ll = np.concatenate([llA, llM])
    assert not stA.any() and not stM.any()
    np.testing.assert_allclose(ll, G["K_ll"], rtol=RTOL)
    # archived logs print 12 significant digits
    np.testing.assert_allclose(ll, G["K_archived"], rtol=5e-12)


@pytest.mark.parametrize("method", ["eigen", "squaring", "squaring_dfma"])
def test_propagator_matches_expm(engine, gold_ll, method):
    from mcdiff_b200.engine import DeviceProblem
    G = gold_ll
    m = METHODS[method]
    pA = DeviceProblem(engine, G["A_counts"], G["A_lags"], True)
    for v, w, Pref in ((G["K_v"][1], G["K_w"][1], G["K2_P"]), (G["rand_v"][0], G["rand_w"][0], G["rand_P0"])):
        _, P, st = pA.loglike(v, w, want_P=True, method=m)
        P = P[0, 0]
        assert not st.any()
        assert np.max(np.abs(P - Pref)) <= RTOL * np.max(np.abs(Pref))
        big = Pref > 1e-6
        assert np.max(np.abs(P - Pref)[big] / Pref[big]) <= RTOL
        np.testing.assert_allclose(P.sum(axis=0), 1.0, atol=1e-12)     # column stochastic
    pC = DeviceProblem(engine, G["C_counts"], G["C_lags"], False)
    _, P, _ = pC.loglike(G["C_v"][0], G["C_w"][0], want_P=True, method=m)
    assert np.max(np.abs(P[0, 0] - G["C_P0"])) <= RTOL * np.max(np.abs(G["C_P0"]))


def test_random_profiles_all_fixtures(engine, gold_ll):
    """pbc 1 lag, pbc 4 lags, nopbc cut (n=60), odd size n=37 with 2 lags (padding path)."""
    from mcdiff_b200.engine import DeviceProblem
    G = gold_ll
    cases = [
        (G["A_counts"], G["A_lags"], True, G["rand_v"], G["rand_w"], G["rand_ll_A"]),
        (G["P_counts"], G["P_lags"], True, G["rand_v"], G["rand_wP"], G["rand_ll_P"]),
        (G["C_counts"], G["C_lags"], False, G["C_v"], G["C_w"], G["C_ll"]),
        (G["S_counts"], G["S_lags"], True, G["S_v"], G["S_w"], G["S_ll"]),
    ]
    for counts, lags, pbc, v, w, ref in cases:
        prob = DeviceProblem(engine, counts, lags, pbc)
        ll, _, st = prob.loglike(v, w)
        assert not st.any()
        np.testing.assert_allclose(ll, ref, rtol=RTOL)
        assert prob.squaring_ok    # incl. lags 5, 12.5 (no progression: two squaring chains)
        for m in ("squaring", "squaring_dfma"):
            ll2, _, st2 = prob.loglike(v, w, method=METHODS[m])
            assert not st2.any()
            np.testing.assert_allclose(ll2, ref, rtol=RTOL)
    # flat start of the 4-lag fixture (-23355.373632376784)
    prob = DeviceProblem(engine, G["P_counts"], G["P_lags"], True)
    ll, _, _ = prob.loglike(np.zeros(100), G["P_flat_w"])
    np.testing.assert_allclose(ll[0], G["P_flat_ll"], rtol=RTOL)


def test_per_profile_lagtimes(engine, gold_ll):
    """lagtimes argument (time-offset moves evaluate the same profile at shifted lags)."""
    from mcdiff_b200.engine import DeviceProblem
    from oracle import mcdiff_oracle as O
    G = gold_ll
    prob = DeviceProblem(engine, G["P_counts"], G["P_lags"], True)
    v, w = G["rand_v"][:3], G["rand_wP"][:3]
    lt = G["P_lags"][None, :] + np.array([[-1.5], [0.0], [2.25]])
    ll, _, _ = prob.loglike(v, w, lagtimes=lt)
    ref = [O.log_like_lag(v[b], w[b], lt[b], G["P_counts"], True) for b in range(3)]
    np.testing.assert_allclose(ll, ref, rtol=RTOL)


def test_nan_profile_reports_status(engine, gold_ll):
    from mcdiff_b200.engine import DeviceProblem
    G = gold_ll
    prob = DeviceProblem(engine, G["P_counts"], G["P_lags"], True)
    v = np.zeros((2, 100))
    v[1, 7] = np.nan
    w = np.tile(G["P_flat_w"], (2, 1))
    ll, _, st = prob.loglike(v, w)
    assert np.isfinite(ll[0]) and st[0] == 0
    assert np.isnan(ll[1]) and st[1] != 0


@pytest.mark.parametrize("method", ["eigen", "squaring", "squaring_dfma"])
@pytest.mark.parametrize("tag", REPLAY_TAGS)
def test_replay_against_reference_trajectories(engine, gold_replay, tag, method):
    """Tape-driven MC: decisions bit-exact vs the UNMODIFIED reference run on the same tape
    (Fourier 10/6, per-bin 4 lags, time-offset + annealing, nopbc cut, spring prior)."""
    from mcdiff_b200.engine import ChainBatch
    case = replay_case(gold_replay, tag)
    g = case["g"]
    prob = _problem(engine, case)
    nch = 3
    ch = ChainBatch(prob, nch)
    ch.set_state(g("v0"), g("w0"), g("vc0") if case["ncosF"] else None, g("wc0") if case["ncosD"] else None,
                 dv=case["dv"], dw=case["dw"], dtimezero=case["dt0"], temp=case["temp"])
    ll0 = ch.init_log_like()
    np.testing.assert_allclose(ll0, float(g("ll0")), rtol=RTOL)
    nmc = case["nmc"]
    dec, llt, samples = ch.run(nmc, _params(case, nmc, method=METHODS[method]), g("tape_u"), g("tape_idx"),
                               record=True, nsamples_cap=nmc // 100)
    dec = dec.cpu().numpy()
    st = ch.host_state()
    for c in range(nch):   # all chains share the tape -> identical trajectories
        assert np.array_equal(dec[c] >> 2, g("move"))
        assert np.array_equal(dec[c] & 1, g("accepted"))
        assert np.array_equal((dec[c] >> 1) & 1, g("accepted_t0"))
    s0 = st["scal"][0]
    np.testing.assert_allclose(s0[0], float(g("ll")), rtol=RTOL)
    np.testing.assert_allclose(s0[1], float(g("timezero")), rtol=0, atol=1e-15)
    np.testing.assert_allclose(s0[2:4], [float(g("dv")), float(g("dw"))], rtol=1e-14)
    np.testing.assert_allclose(s0[5], float(g("temp")), rtol=1e-14)
    np.testing.assert_allclose(st["v"][0], g("v"), rtol=0, atol=1e-13)
    np.testing.assert_allclose(st["w"][0], g("w"), rtol=0, atol=1e-13)
    assert list(st["counters"][0][:3]) == [int(g("naccv")), int(g("naccw")), int(g("nacct0"))]
    assert st["counters"][0][6] == 0
    if case["ncosF"]:
        assert np.array_equal(st["naccv_coeff"][0], g("naccv_coeff"))
    if case["ncosD"]:
        assert np.array_equal(st["naccw_coeff"][0], g("naccw_coeff"))
    # Logger rows (lib/log.py:35-58): rows 1.. of the reference logger
    smp = samples.cpu().numpy()[0]
    np.testing.assert_allclose(smp[:, 0], g("log_ll")[1:], rtol=RTOL)
    np.testing.assert_allclose(smp[:, 2], g("log_dv")[1:], rtol=1e-14)
    nv = g("log_pv").shape[1]
    np.testing.assert_allclose(smp[:, 6:6 + nv], g("log_pv")[1:], rtol=0, atol=1e-13)
    np.testing.assert_allclose(smp[:, 6 + nv:], g("log_pw")[1:], rtol=0, atol=1e-13)


@pytest.mark.parametrize("method", ["eigen", "squaring", "squaring_dfma"])
def test_philox_run_equals_oracle_replay(engine, gold_ll, method):
    """Production mode (Philox4x32-10 on device): the trajectory of each chain equals the
    oracle replayed on the tape the host mirror of the RNG derives; chains are split over two
    launches and over a chain_offset to cover resumption and multi-GPU sharding."""
    from mcdiff_b200.engine import DeviceProblem, ChainBatch
    from mcdiff_b200 import philox
    from oracle import mcdiff_oracle as O
    G = gold_ll
    n = 100
    vb, wb = O.cos_basis_center(n, 6), O.cos_basis_border(n, n, 4)
    prob = DeviceProblem(engine, G["P_counts"], G["P_lags"], True, v_basis=vb, w_basis=wb)
    nch, nsteps, seed, off = 4, 160, 20240607, 1000
    wc0 = np.zeros(4)
    wc0[0] = G["P_flat_w"][0]
    ch = ChainBatch(prob, nch)
    ch.set_state(np.zeros(n), wb @ wc0, np.zeros(6), wc0, dv=0.3, dw=0.3)
    ch.init_log_like()
    case = dict(nmc_update=50.0, dtemp=0.0, lagtimes=G["P_lags"], move_timezero=False)
    mp = _params(case, nsteps, use_tape=0, seed=seed, chain_offset=off, method=METHODS[method])
    d1, _, _ = ch.run(100, mp, record=True)
    d2, _, _ = ch.run(60, mp, record=True)
    dec = np.concatenate([d1.cpu().numpy(), d2.cpu().numpy()], axis=1)
    st = ch.host_state()
    cfg = O.ChainConfig(G["P_counts"], G["P_lags"], True, v_basis=vb, w_basis=wb, dv=0.3, dw=0.3, temp=1.0,
                        nmc=nsteps, num_mc_update=50.0)
    for c in range(nch):
        tu, ti = philox.chain_tape(seed, off + c, 0, nsteps, n, n, 6, 4)
        out = O.replay(cfg, np.zeros(n), wb @ wc0, tu, ti, v_coeff0=np.zeros(6), w_coeff0=wc0)
        assert np.array_equal(dec[c] & 1, out["accepted"]), "chain %d" % c
        assert np.array_equal(dec[c] >> 2, out["move"])
        np.testing.assert_allclose(st["scal"][c][0], out["ll"], rtol=RTOL)
        np.testing.assert_allclose(st["vcoef"][c], out["v_coeff"], atol=1e-13)
        np.testing.assert_allclose(st["scal"][c][2:4], [out["dv"], out["dw"]], rtol=1e-13)
    # different chains really are different streams
    assert not np.array_equal(dec[0], dec[1])


@pytest.mark.parametrize("method", ["eigen", "squaring", "squaring_dfma"])
def test_full_size_properties(engine, method):
    """BASELINE shape (100 bins x 4 lags, 1024 chains): size-independent checks.
    (a) log L of the final state recomputed by the cold batched path equals the running
    log L each chain carried through warm-started solves; (b) propagators of final states are
    column-stochastic and satisfy detailed balance P_ij pi_j = P_ji pi_i."""
    from mcdiff_b200.engine import DeviceProblem, ChainBatch
    from oracle import mcdiff_oracle as O
    counts, lags, edges, F, w_true = O.synthetic_counts()
    n = 100
    prob = DeviceProblem(engine, counts, lags, True)
    nch = 1024
    w0 = np.zeros(n) + np.log(1.0 / 0.25)
    ch = ChainBatch(prob, nch)
    ch.set_state(np.zeros(n), w0, dv=0.05, dw=0.05)
    ll0 = ch.init_log_like()
    ref0 = O.log_like_lag(np.zeros(n), w0, lags, counts, True)
    np.testing.assert_allclose(ll0, ref0, rtol=RTOL)
    case = dict(nmc_update=100.0, dtemp=0.0, lagtimes=lags, move_timezero=False)
    mp = _params(case, 0, use_tape=0, seed=1, refresh=512, method=METHODS[method])
    ch.run(300, mp)
    st = ch.host_state()
    assert not st["counters"][:, 6].any()
    ll_cold, _, stat = prob.loglike(st["v"], st["w"])
    assert not stat.any()
    np.testing.assert_allclose(st["scal"][:, 0], ll_cold, rtol=1e-11)
    assert np.all(st["scal"][:, 0] > ll0)          # every chain climbed from the flat start
    acc = (st["counters"][:, 0] + st["counters"][:, 1]) / 300.0
    assert 0.02 < acc.mean() < 0.9
    # oracle on a few final states
    for c in (0, 511, 1023):
        np.testing.assert_allclose(st["scal"][c, 0], O.log_like_lag(st["v"][c], st["w"][c], lags, counts, True),
                                   rtol=RTOL)
    _, P, _ = prob.loglike(st["v"][:4], st["w"][:4], want_P=True)
    for c in range(4):
        pi = np.exp(-st["v"][c])
        for l in range(4):
            np.testing.assert_allclose(P[c, l].sum(axis=0), 1.0, atol=1e-12)
            flux = P[c, l] * pi[None, :]
            assert np.max(np.abs(flux - flux.T)) <= 1e-12 * np.max(flux)


def test_radial_likelihood(engine, gold_radial):
    from mcdiff_b200.engine import rad_loglike
    Rd = gold_radial
    ll, P, st = rad_loglike(engine, Rd["v"], Rd["w"], True, Rd["lags"], Rd["counts"], Rd["table"], Rd["zeros"],
                            Rd["wrad"], want_P=True)
    assert not st.any()
    np.testing.assert_allclose(ll, Rd["ll"], rtol=RTOL)
    assert np.max(np.abs(P[0, 0] - Rd["P0"])) <= RTOL * np.max(np.abs(Rd["P0"]))


def test_large_n_generic_path(engine):
    """n = 200, 8 lags (BASELINE config 3 shape): the work matrices do not fit in shared
    memory.  The eigen engine runs on global work matrices, the squaring engine blocked (2 x 2
    blocks through shared-memory panels); both must give the oracle's numbers."""
    from mcdiff_b200.engine import DeviceProblem
    from oracle import mcdiff_oracle as O
    lags = (5.0, 10.0, 15.0, 20.0, 25.0, 30.0, 35.0, 40.0)
    counts, lags, edges, F, w_true = O.synthetic_counts(n=200, lags=lags, seed=3, total=2.0e5)
    prob = DeviceProblem(engine, counts, lags, True)
    assert not prob.in_smem and prob.squaring_ok
    rng = np.random.default_rng(1)
    v = F[None, :] + rng.normal(0, 0.2, (3, 200))
    w = w_true[None, :] + rng.normal(0, 0.2, (3, 200))
    ref = [O.log_like_lag(v[b], w[b], lags, counts, True) for b in range(3)]
    for method in ("eigen", "squaring"):
        ll, P, st = prob.loglike(v, w, want_P=True, method=METHODS[method])
        assert not st.any()
        np.testing.assert_allclose(ll, ref, rtol=RTOL)
        for l in (0, 7):
            Pref = O.propagator(v[1], w[1], lags[l], True)
            assert np.max(np.abs(P[1, l] - Pref)) < 1e-12


@pytest.mark.parametrize("n,lags,pbc", [(120, (5.0, 10.0, 15.0), True), (200, (5.0, 10.0, 20.0, 40.0), True),
                                        (136, (3.0, 7.0, 10.0), False), (106, (10.0,), True),
                                        (208, (5.0, 10.0), True), (168, (4.0, 4.0, 9.0), True),
                                        (250, (5.0, 10.0), True), (330, (4.0, 8.0, 12.0), False),
                                        (508, (6.0, 12.0), True)])
def test_blocked_squaring_engine(engine, n, lags, pbc):
    """104 < n <= 208 on the tensor-core squaring engine: fast and general lag plans, non-periodic,
    single lag, duplicate lag, the largest size; more profiles than CTAs of one wave are not needed,
    but several per CTA exercise the workspace reuse."""
    from mcdiff_b200.engine import DeviceProblem
    from oracle import mcdiff_oracle as O
    lags = np.array(lags)
    rng = np.random.default_rng(3)
    if pbc:
        counts = O.synthetic_counts(n, lags, seed=3, total=2.0e5, dz=1.0)[0]
    else:
        counts = rng.integers(0, 50, size=(len(lags), n, n)).astype(np.int32)
        ii, jj = np.indices((n, n))
        counts[:, np.abs(ii - jj) > 12] = 0
    dim_w = n if pbc else n - 1
    B = 5
    v = rng.normal(0, 0.5, (B, n))
    w = rng.normal(-1.0, 0.3, (B, dim_w))
    prob = DeviceProblem(engine, counts, lags, pbc)
    assert prob.squaring_ok and not prob.in_smem
    ll, P, st = prob.loglike(v, w, want_P=True, method=METHODS["squaring"])
    assert not st.any()
    for b in range(B):
        np.testing.assert_allclose(ll[b], O.log_like_lag(v[b], w[b], lags, counts, pbc), rtol=1e-12)
    for l, lt in enumerate(lags):
        Pref = O.propagator(v[2], w[2], lt, pbc)
        assert np.max(np.abs(P[2, l] - Pref)) < 1e-13
    with pytest.raises(Exception):
        prob.loglike(v, w, method=METHODS["squaring_dfma"])


@pytest.mark.parametrize("cluster", ["0", "auto"])
def test_blocked_engine_mc_matches_oracle(engine, monkeypatch, cluster):
    """MC loop on the blocked engine (n = 120, cosine model, time-offset moves, a lag set that
    needs a parked operand), device Philox: every chain's decisions equal the oracle replay.
    cluster "0": one SM per chain; "auto": a batch smaller than a wave runs on two-CTA clusters (mcd_mc_run)."""
    if cluster == "0":
        monkeypatch.setenv("MCDIFF_B200_CLUSTER", "0")
    else:
        monkeypatch.delenv("MCDIFF_B200_CLUSTER", raising=False)
    from mcdiff_b200.engine import DeviceProblem, ChainBatch
    from mcdiff_b200 import philox
    from oracle import mcdiff_oracle as O
    n, lags = 120, np.array([5.0, 10.0, 20.0])
    counts = O.synthetic_counts(n, lags, seed=5, total=2.0e5, dz=1.0)[0]
    vb, wb = O.cos_basis_center(n, 5), O.cos_basis_border(n, n, 4)
    wc0 = np.zeros(4)
    wc0[0] = -1.0
    w0 = wb @ wc0
    nch, nsteps, seed, off = 3, 40, 77, 10
    prob = DeviceProblem(engine, counts, lags, True, v_basis=vb, w_basis=wb)
    ch = ChainBatch(prob, nch)
    ch.set_state(np.zeros(n), w0, np.zeros(5), wc0, dv=0.2, dw=0.2, dtimezero=0.3)
    ch.init_log_like()
    case = dict(nmc_update=10.0, dtemp=0.0, lagtimes=lags, move_timezero=True)
    mp = _params(case, nsteps, use_tape=0, seed=seed, chain_offset=off, method=METHODS["squaring"])
    d1, _, _ = ch.run(25, mp, record=True)
    d2, _, _ = ch.run(15, mp, record=True)
    dec = np.concatenate([d1.cpu().numpy(), d2.cpu().numpy()], axis=1)
    st = ch.host_state()
    assert not st["counters"][:, 6].any()
    cfg = O.ChainConfig(counts, lags, True, v_basis=vb, w_basis=wb, dv=0.2, dw=0.2, dtimezero=0.3, temp=1.0,
                        nmc=nsteps, num_mc_update=10.0, move_timezero=True)
    for c in range(nch):
        tu, ti = philox.chain_tape(seed, off + c, 0, nsteps, n, n, 5, 4)
        out = O.replay(cfg, np.zeros(n), w0, tu, ti, v_coeff0=np.zeros(5), w_coeff0=wc0)
        assert np.array_equal(dec[c] & 1, out["accepted"]), "chain %d" % c
        assert np.array_equal((dec[c] >> 1) & 1, out["accepted_t0"])
        np.testing.assert_allclose(st["scal"][c][0], out["ll"], rtol=RTOL)
        np.testing.assert_allclose(st["scal"][c][1], out["timezero"], rtol=1e-13)


def test_engines_agree_step_by_step(engine):
    """Same Philox streams through both propagator engines: identical decisions for every
    chain and step (each engine's log L error is orders of magnitude below the margins)."""
    from mcdiff_b200.engine import DeviceProblem, ChainBatch
    from oracle import mcdiff_oracle as O
    counts, lags, edges, F, w_true = O.synthetic_counts()
    n = 100
    prob = DeviceProblem(engine, counts, lags, True)
    w0 = np.zeros(n) + np.log(1.0 / 0.25)
    out = {}
    for name, m in METHODS.items():
        ch = ChainBatch(prob, 64)
        ch.set_state(np.zeros(n), w0, dv=0.05, dw=0.05)
        ch.init_log_like()
        case = dict(nmc_update=100.0, dtemp=0.0, lagtimes=lags, move_timezero=False)
        dec, llt, _ = ch.run(200, _params(case, 0, use_tape=0, seed=7, method=m), record=True)
        out[name] = (dec.cpu().numpy(), llt.cpu().numpy(), ch.host_state())
    for other in ("squaring", "squaring_dfma"):
        assert np.array_equal(out["eigen"][0], out[other][0])
        np.testing.assert_allclose(out["eigen"][1], out[other][1], rtol=1e-12)
        np.testing.assert_allclose(out["eigen"][2]["v"], out[other][2]["v"], atol=1e-13)


def test_squaring_general_lag_sets(engine, gold_ll):
    """Lag sets that are no arithmetic progression, unsorted, with duplicates (replica files of the
    same lag), and the shipped 11-lag pattern: the squaring engine plans chains / parked operands."""
    from mcdiff_b200.engine import DeviceProblem
    from oracle import mcdiff_oracle as O
    G = gold_ll
    v, w = G["rand_v"][:2], G["rand_wP"][:2]
    for lags in ([10.0, 20.0, 35.0, 40.0], [10.0, 20.0, 20.0, 40.0], [40.0, 10.0, 30.0, 20.0], [7.5, 20.0, 30.0, 45.0]):
        prob = DeviceProblem(engine, G["P_counts"], lags, True)
        assert prob.squaring_ok
        ll, P, st = prob.loglike(v, w, method=METHODS["squaring"], want_P=True)
        assert not st.any()
        ref = [O.log_like_lag(v[b], w[b], lags, G["P_counts"], True) for b in range(2)]
        np.testing.assert_allclose(ll, ref, rtol=RTOL)
        Pref = O.propagator(v[0], w[0], lags[2], True)
        assert np.max(np.abs(P[0, 2] - Pref)) <= RTOL * np.max(Pref)
    lags11 = [1.0, 2.0, 3.0, 4.0, 5.0, 10.0, 15.0, 20.0, 30.0, 40.0, 50.0]
    counts11 = np.concatenate([G["P_counts"], G["P_counts"], G["P_counts"][:3]])
    prob = DeviceProblem(engine, counts11, lags11, True)
    ll, _, st = prob.loglike(v, w, method=METHODS["squaring"])
    ref = [O.log_like_lag(v[b], w[b], lags11, counts11, True) for b in range(2)]
    np.testing.assert_allclose(ll, ref, rtol=RTOL)


def test_squaring_rejects_unsupported(engine):
    """A lag set without a LagPlan (nine incommensurate differences): SQUARING is refused, AUTO falls back to the
    eigen engine (here on global work matrices, n = 240)."""
    from mcdiff_b200.engine import DeviceProblem
    from mcdiff_b200._capi import McdError
    from oracle import mcdiff_oracle as O
    rng = np.random.default_rng(0)
    lags = np.cumsum([1.0, 1.1, 1.23, 1.37, 1.51, 1.69, 1.83, 1.97, 2.11])
    counts = rng.integers(0, 5, size=(len(lags), 240, 240))
    prob = DeviceProblem(engine, counts, lags, True)
    assert not prob.squaring_ok and not prob.in_smem
    with pytest.raises(McdError):
        prob.loglike(np.zeros(240), np.zeros(240), method=METHODS["squaring"])
    ll, _, st = prob.loglike(np.zeros(240), np.zeros(240))
    assert not st.any()
    np.testing.assert_allclose(ll[0], O.log_like_lag(np.zeros(240), np.zeros(240), lags, counts, True), rtol=RTOL)


@pytest.mark.parametrize("lags", [(5.0,), (5.0, 10.0, 15.0), (4.0, 6.0, 10.0),
                                  tuple(np.cumsum([1.0, 1.1, 1.23, 1.37, 1.51, 1.69, 1.83, 1.97, 2.11]))])
def test_radial_likelihood_lag_sets(engine, lags):
    """Radial likelihood against the oracle for a single lag, an arithmetic progression (fast
    LagPlan), a general plan with a parked operand, and a lag set without a plan (nine
    incommensurate differences -> eigen fallback)."""
    from mcdiff_b200.engine import rad_loglike
    from oracle import mcdiff_oracle as O
    n, dim_rad, lmax = 24, 15, 9
    rng = np.random.default_rng(17)
    i = np.arange(n)
    v = 1.2 * np.cos(2 * np.pi * (i + 0.5) / n)
    w = np.log(0.5 + 0.2 * np.cos(2 * np.pi * (i + 1.0) / n))
    lags = np.array(lags)
    zeros, table = O.bessel_tables(lmax, dim_rad)
    rate = O.rate_matrix(v, w, True)
    wrad0 = np.full(n, np.log(0.4))
    counts = np.zeros((len(lags), dim_rad, n, n))
    for l, lt in enumerate(lags):
        counts[l] = rng.poisson(300.0 * np.clip(O.radial_propagator(rate, wrad0, lt, zeros, table), 0.0, None))
    wrad = wrad0[None, :] + rng.normal(0, 0.2, (3, n))
    ll, P, st = rad_loglike(engine, v, w, True, lags, counts, table, zeros, wrad, want_P=True)
    assert not st.any()
    for b in range(3):
        np.testing.assert_allclose(ll[b], O.rad_log_like_lag(rate, wrad[b], lags, counts, zeros, table), rtol=RTOL)
    Pref = O.radial_propagator(rate, wrad[1], lags[-1], zeros, table)
    assert np.max(np.abs(P[1, -1] - Pref)) <= 1e-12 * np.max(np.abs(Pref))


@pytest.mark.parametrize("tag", ["radbin", "radcos"])
def test_radial_mc_replay_against_reference(engine, tag):
    """Radial Monte-Carlo loop (per-bin and cosine wrad moves): decisions bit-exact vs the
    unmodified reference's mcmove_diffusion_radial driven by the same tape."""
    import os
    from mcdiff_b200.engine import RadialChains
    from oracle import mcdiff_oracle as O
    RR = np.load(os.path.join(os.path.dirname(__file__), "golden", "radial_replay_golden.npz"))
    g = lambda k: RR[tag + "/" + k]  # noqa: E731
    n = len(g("v"))
    ncos = len(g("wc0"))
    basis = O.cos_basis_center(n, ncos) if ncos else None
    ch = RadialChains(engine, g("v"), g("w"), True, g("lags"), g("counts"), g("table"), g("zeros"), 2, wrad_basis=basis)
    ch.set_state(g("wrad0"), g("wc0") if ncos else None, dwrad=0.4, temp=1.0)
    ll0 = ch.init_log_like()
    np.testing.assert_allclose(ll0, float(g("ll0")), rtol=RTOL)
    nmc = len(g("accepted"))
    dec, llt, smp = ch.run(nmc, num_mc_update=40.0, tape_u=g("tape_u"), tape_idx=g("tape_idx"), record=True,
                           nsamples_cap=nmc // 100)
    dec = dec.cpu().numpy()
    st = ch.host_state()
    for c in range(2):
        assert np.array_equal(dec[c] & 1, g("accepted"))
    np.testing.assert_allclose(st["scal"][0, 0], float(g("ll")), rtol=RTOL)
    np.testing.assert_allclose(st["scal"][0, 1], float(g("dwrad")), rtol=1e-13)
    np.testing.assert_allclose(st["wrad"][0], g("wrad"), atol=1e-13)
    assert st["counters"][0, 0] == int(g("naccwrad")) and st["counters"][0, 3] == 0
    if ncos:
        np.testing.assert_allclose(st["wcoef"][0], g("wc"), atol=1e-13)
    s0 = smp.cpu().numpy()[0]
    np.testing.assert_allclose(s0[:, 0], g("log_ll")[1:], rtol=RTOL)
    np.testing.assert_allclose(s0[:, 3:], g("log_wrad")[1:], atol=1e-13)


def test_pull_likelihood_and_propagator(engine, gold_pull):
    """--pull (lib/utils.py:133-183): non-symmetrisable generator on the general squaring engine, against
    values the unmodified reference produced; eigen / DFMA are refused for such a problem."""
    from mcdiff_b200.engine import DeviceProblem
    from mcdiff_b200._capi import McdError
    G = gold_pull
    for b in range(4):
        prob = DeviceProblem(engine, G["counts"], G["lags"], True, pull=float(G["ll/dF"][b]))
        ll, P, st = prob.loglike(G["ll/v"][b:b + 1], G["ll/w"][b:b + 1], want_P=True)
        assert not st.any()
        np.testing.assert_allclose(ll[0], G["ll/ll"][b], rtol=1e-12)
        for l in range(4):
            np.testing.assert_allclose(P[0, l].sum(axis=0), 1.0, atol=1e-12)
        if b == 0:
            assert np.max(np.abs(P[0, 0] - G["P0"])) < 1e-13
            with pytest.raises(McdError):
                prob.loglike(G["ll/v"][:1], G["ll/w"][:1], method=METHODS["eigen"])
    with pytest.raises(NotImplementedError):
        DeviceProblem(engine, G["counts"], G["lags"], False, pull=0.1)


@pytest.mark.parametrize("tag", PULL_TAGS)
def test_pull_replay_against_reference_trajectories(engine, gold_pull, tag):
    """Full --pull trajectories of the reference (with and without time-offset moves): bit-exact decisions."""
    from mcdiff_b200.engine import DeviceProblem, ChainBatch
    case = replay_case(gold_pull, tag)
    g = case["g"]
    prob = DeviceProblem(engine, case["counts"], case["lagtimes"], True, v_basis=case["v_basis"],
                         w_basis=case["w_basis"], pull=case["pull"])
    ch = ChainBatch(prob, 2)
    ch.set_state(g("v0"), g("w0"), g("vc0") if case["ncosF"] else None, g("wc0") if case["ncosD"] else None,
                 dv=case["dv"], dw=case["dw"], dtimezero=case["dt0"], temp=case["temp"])
    ll0 = ch.init_log_like()
    np.testing.assert_allclose(ll0, float(g("ll0")), rtol=1e-12)
    nsteps = case["nmc"]
    mp = _params(case, nsteps, use_tape=1, method=METHODS["squaring"])
    dec, llt, _ = ch.run(nsteps, mp, tape_u=g("tape_u"), tape_idx=g("tape_idx"), record=True)
    dec = dec.cpu().numpy()
    for c in range(2):
        assert np.array_equal(dec[c] >> 2, g("move"))
        assert np.array_equal(dec[c] & 1, g("accepted"))
        assert np.array_equal((dec[c] >> 1) & 1, g("accepted_t0"))
    st = ch.host_state()
    np.testing.assert_allclose(st["scal"][0][0], float(g("ll")), rtol=1e-11)
    np.testing.assert_allclose(st["scal"][0][1], float(g("timezero")), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(st["v"][0], g("v"), atol=1e-12)
    assert not st["counters"][:, 6].any()


@pytest.mark.parametrize("n,dim_rad,lmax", [(8, 7, 5), (30, 60, 12), (52, 20, 33)])
def test_radial_shapes(engine, n, dim_rad, lmax):
    """Radial likelihood at the edges of the batched route: a tiny problem (Bessel sums fall back to the vector
    pipe), more than 48 radial bins (two accumulator groups), more Bessel terms than fit one k-step multiple."""
    from mcdiff_b200.engine import rad_loglike
    from oracle import mcdiff_oracle as O
    rng = np.random.default_rng(n)
    i = np.arange(n)
    v = 1.0 * np.cos(2 * np.pi * (i + 0.5) / n)
    w = np.log(0.5 + 0.2 * np.cos(2 * np.pi * (i + 1.0) / n))
    lags = np.array([4.0, 8.0])
    zeros, table = O.bessel_tables(lmax, dim_rad)
    rate = O.rate_matrix(v, w, True)
    wrad0 = np.full(n, np.log(0.4))
    counts = np.zeros((2, dim_rad, n, n))
    for l, lt in enumerate(lags):
        counts[l] = rng.poisson(200.0 * np.clip(O.radial_propagator(rate, wrad0, lt, zeros, table), 0.0, None))
    wrad = wrad0[None, :] + rng.normal(0, 0.2, (2, n))
    ll, P, st = rad_loglike(engine, v, w, True, lags, counts, table, zeros, wrad, want_P=True)
    assert not st.any()
    for b in range(2):
        np.testing.assert_allclose(ll[b], O.rad_log_like_lag(rate, wrad[b], lags, counts, zeros, table), rtol=RTOL)
    Pref = O.radial_propagator(rate, wrad[0], lags[1], zeros, table)
    assert np.max(np.abs(P[0, 1] - Pref)) <= 1e-12 * np.max(np.abs(Pref))


@pytest.mark.parametrize("pbc", [True, False])
def test_small_and_odd_sizes(engine, pbc):
    """n = 3 ... 104 including sizes that are not multiples of 8 (padding), fewer rows than threads per row-slot
    tables, band wider than the matrix: squaring engines to 1e-11 (componentwise accurate), eigen to 1e-8 on these
    unphysical random counts (absolute, not componentwise, accuracy)."""
    from mcdiff_b200.engine import DeviceProblem
    from oracle import mcdiff_oracle as O
    rng = np.random.default_rng(5 + int(pbc))
    for n in (3, 4, 5, 8, 9, 12, 17, 24, 33, 49, 50, 64, 97, 104):
        lags = np.array([2.0, 4.0, 6.0]) if n % 2 else np.array([3.0, 5.0])
        counts = rng.integers(0, 40, size=(len(lags), n, n)).astype(np.int32)
        dim_w = n if pbc else n - 1
        v = rng.normal(0, 0.5, (2, n))
        w = rng.normal(-0.5, 0.4, (2, dim_w))
        prob = DeviceProblem(engine, counts, lags, pbc)
        ref = np.array([O.log_like_lag(v[b], w[b], lags, counts, pbc) for b in range(2)])
        for method, tol in (("squaring", 1e-11), ("squaring_dfma", 1e-11), ("eigen", 1e-8)):
            ll, _, st = prob.loglike(v, w, method=METHODS[method])
            assert not st.any(), (n, method)
            np.testing.assert_allclose(ll, ref, rtol=tol, err_msg="n=%d %s" % (n, method))
        if pbc and n >= 8:
            probp = DeviceProblem(engine, counts, lags, True, pull=0.15)
            llp, _, st = probp.loglike(v, w)
            refp = [O.log_like_lag(v[b], w[b], lags, counts, True, pull=0.15) for b in range(2)]
            np.testing.assert_allclose(llp, refp, rtol=1e-11, err_msg="pull n=%d" % n)


def test_pull_time_offset_through_the_likelihood_seam(engine, gold_pull):
    """mcd_loglike_batch on a --pull problem with per-profile lag times = lag times + offset (what
    mcmove_timezero passes, lib/MCState.py:207-210); anything else is flagged, not silently computed."""
    from mcdiff_b200.engine import DeviceProblem
    from oracle import mcdiff_oracle as O
    G = gold_pull
    dF = float(G["ll/dF"][1])
    prob = DeviceProblem(engine, G["counts"], G["lags"], True, pull=dF)
    v, w = G["ll/v"][:2], G["ll/w"][:2]
    lt = np.stack([G["lags"] + 0.37, G["lags"] - 1.25])
    ll, _, st = prob.loglike(v, w, lagtimes=lt)
    assert not st.any()
    for b in range(2):
        np.testing.assert_allclose(ll[b], O.log_like_lag(v[b], w[b], lt[b], G["counts"], True, pull=dF), rtol=1e-12)
    bad = lt.copy()
    bad[1, 2] += 0.5
    ll, _, st = prob.loglike(v, w, lagtimes=bad)
    assert st[0] == 0 and st[1] == 4 and np.isnan(ll[1]) and np.isfinite(ll[0])


@pytest.mark.parametrize("n,pbc", [(48, True), (24, True), (40, False), (56, True)])
def test_small_profiles_run_on_the_128_thread_kernel(engine, n, pbc):
    """np <= 56: mcd_mc_run dispatches to the 128-thread build of the chain body (three CTAs per SM); device Philox
    trajectories must equal the oracle replay exactly like the 384-thread kernel's."""
    from mcdiff_b200.engine import DeviceProblem, ChainBatch
    from mcdiff_b200 import philox
    from oracle import mcdiff_oracle as O
    lags = np.array([5.0, 10.0, 15.0])
    rng = np.random.default_rng(n)
    if pbc:
        counts = O.synthetic_counts(n, lags, seed=8, total=1.0e5, dz=1.0)[0]
    else:
        counts = rng.integers(0, 60, size=(3, n, n)).astype(np.int32)
        ii, jj = np.indices((n, n))
        counts[:, np.abs(ii - jj) > 6] = 0
    dim_w = n if pbc else n - 1
    vb, wb = O.cos_basis_center(n, 5), O.cos_basis_border(dim_w, n, 4)
    wc0 = np.zeros(4)
    wc0[0] = -1.0
    w0 = wb @ wc0
    nch, nsteps, seed = 5, 50, 31
    prob = DeviceProblem(engine, counts, lags, pbc, v_basis=vb, w_basis=wb)
    ch = ChainBatch(prob, nch)
    ch.set_state(np.zeros(n), w0, np.zeros(5), wc0, dv=0.3, dw=0.3, dtimezero=0.3)
    ch.init_log_like()
    case = dict(nmc_update=10.0, dtemp=0.0, lagtimes=lags, move_timezero=True)
    mp = _params(case, nsteps, use_tape=0, seed=seed, method=METHODS["squaring"])
    d1, _, _ = ch.run(30, mp, record=True)
    d2, _, _ = ch.run(20, mp, record=True)
    dec = np.concatenate([d1.cpu().numpy(), d2.cpu().numpy()], axis=1)
    st = ch.host_state()
    assert not st["counters"][:, 6].any()
    cfg = O.ChainConfig(counts, lags, pbc, v_basis=vb, w_basis=wb, dv=0.3, dw=0.3, dtimezero=0.3, temp=1.0, nmc=nsteps,
                        num_mc_update=10.0, move_timezero=True)
    for c in range(nch):
        tu, ti = philox.chain_tape(seed, c, 0, nsteps, n, dim_w, 5, 4)
        out = O.replay(cfg, np.zeros(n), w0, tu, ti, v_coeff0=np.zeros(5), w_coeff0=wc0)
        assert np.array_equal(dec[c] & 1, out["accepted"]), "chain %d" % c
        assert np.array_equal((dec[c] >> 1) & 1, out["accepted_t0"])
        np.testing.assert_allclose(st["scal"][c][0], out["ll"], rtol=RTOL)
        np.testing.assert_allclose(st["scal"][c][2:4], [out["dv"], out["dw"]], rtol=1e-13)


def test_blocked_engine_three_blocks_mc(engine):
    """n = 250 (3 x 3 blocks of 88 bins): MC loop with time-offset moves on device Philox streams against the oracle."""
    from mcdiff_b200.engine import DeviceProblem, ChainBatch
    from mcdiff_b200 import philox
    from oracle import mcdiff_oracle as O
    n, lags = 250, np.array([5.0, 10.0])
    counts = O.synthetic_counts(n, lags, seed=6, total=2.0e5, dz=1.0)[0]
    vb, wb = O.cos_basis_center(n, 5), O.cos_basis_border(n, n, 4)
    wc0 = np.zeros(4)
    wc0[0] = -1.0
    w0 = wb @ wc0
    nch, nsteps, seed = 2, 24, 17
    prob = DeviceProblem(engine, counts, lags, True, v_basis=vb, w_basis=wb)
    assert prob.squaring_ok and not prob.in_smem
    ch = ChainBatch(prob, nch)
    ch.set_state(np.zeros(n), w0, np.zeros(5), wc0, dv=0.2, dw=0.2, dtimezero=0.3)
    ch.init_log_like()
    case = dict(nmc_update=10.0, dtemp=0.0, lagtimes=lags, move_timezero=True)
    mp = _params(case, nsteps, use_tape=0, seed=seed, method=METHODS["squaring"])
    dec, _, _ = ch.run(nsteps, mp, record=True)
    dec = dec.cpu().numpy()
    st = ch.host_state()
    cfg = O.ChainConfig(counts, lags, True, v_basis=vb, w_basis=wb, dv=0.2, dw=0.2, dtimezero=0.3, temp=1.0, nmc=nsteps,
                        num_mc_update=10.0, move_timezero=True)
    for c in range(nch):
        tu, ti = philox.chain_tape(seed, c, 0, nsteps, n, n, 5, 4)
        out = O.replay(cfg, np.zeros(n), w0, tu, ti, v_coeff0=np.zeros(5), w_coeff0=wc0)
        assert np.array_equal(dec[c] & 1, out["accepted"]), "chain %d" % c
        assert np.array_equal((dec[c] >> 1) & 1, out["accepted_t0"])
        np.testing.assert_allclose(st["scal"][c][0], out["ll"], rtol=RTOL)


@pytest.mark.parametrize("n", [100, 200])
def test_time_offset_through_the_likelihood_seam_on_squaring_engines(engine, n):
    """mcd_loglike_batch with explicit SQUARING and per-profile lag times = lag times + offset (the time-offset
    trial of lib/MCState.py:207-210) on the shared-memory and the blocked engine; other lag times are flagged."""
    from mcdiff_b200.engine import DeviceProblem
    from oracle import mcdiff_oracle as O
    lags = np.array([5.0, 10.0, 15.0])
    counts, _, _, F, w_true = O.synthetic_counts(n=n, lags=lags, seed=3, total=2.0e5)
    prob = DeviceProblem(engine, counts, lags, True)
    rng = np.random.default_rng(n)
    v = F[None, :] + rng.normal(0, 0.2, (2, n))
    w = w_true[None, :] + rng.normal(0, 0.2, (2, n))
    lt = np.stack([lags + 0.41, lags - 0.9])
    ll, _, st = prob.loglike(v, w, lagtimes=lt, method=METHODS["squaring"])
    assert not st.any()
    for b in range(2):
        np.testing.assert_allclose(ll[b], O.log_like_lag(v[b], w[b], lt[b], counts, True), rtol=1e-12)
    bad = lt.copy()
    bad[0, 1] *= 1.01
    ll, _, st = prob.loglike(v, w, lagtimes=bad, method=METHODS["squaring"])
    assert st[0] == 4 and st[1] == 0 and np.isnan(ll[0]) and np.isfinite(ll[1])
    ll_e, _, st = prob.loglike(v, w, lagtimes=bad)          # AUTO: eigen engine, any lag times
    assert not st.any()
    np.testing.assert_allclose(ll_e[0], O.log_like_lag(v[0], w[0], bad[0], counts, True), rtol=RTOL)


def test_randomised_configurations(engine, capsys):
    """profiles/fuzz_parity.py: random size / lag set / boundary / model / spring / time offset / pull, batched
    likelihood + propagator against the oracle and a short device-Philox run against the oracle replay."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(root, "profiles", "fuzz_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    failures = mod.run(16, 20241021, eng=engine)
    out = capsys.readouterr().out
    assert failures == 0, out

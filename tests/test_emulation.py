"""CPU emulation of the CUDA chain body (same source, HostExec policy) against the
reference fixtures.  This exercises, without a GPU, everything but the hardware: tile maps,
the Jacobi schedule, the scaling-and-squaring chain, Metropolis, adaptation, sampling."""
import numpy as np
import pytest

from mcdiff_b200 import _capi as K
from tests.emul import harness as H
from tests.helpers import replay_case

EIGEN, SQUARING, SQUARING_DFMA = 1, 2, 3


@pytest.mark.parametrize("method", [EIGEN, SQUARING, SQUARING_DFMA])
def test_loglike_and_propagator(gold_ll, method):
    G = gold_ll
    pA = H.HostProblem(G["A_counts"], G["A_lags"], True)
    ll, P, st, _ = H.loglike_batch(pA, G["K_v"][:2], G["K_w"][:2], want_P=True, method=method)
    assert not st.any()
    np.testing.assert_allclose(ll, G["K_ll"][:2], rtol=1e-12)
    assert np.max(np.abs(P[1, 0] - G["K2_P"])) < 1e-13
    np.testing.assert_allclose(P[1, 0].sum(axis=0), 1.0, atol=1e-12)
    pP = H.HostProblem(G["P_counts"], G["P_lags"], True)
    ll, _, st, _ = H.loglike_batch(pP, G["rand_v"][:3], G["rand_wP"][:3], method=method)
    np.testing.assert_allclose(ll, G["rand_ll_P"][:3], rtol=1e-12)
    pC = H.HostProblem(G["C_counts"], G["C_lags"], False)
    ll, P, st, _ = H.loglike_batch(pC, G["C_v"][:2], G["C_w"][:2], want_P=True, method=method)
    np.testing.assert_allclose(ll, G["C_ll"][:2], rtol=1e-12)
    assert np.max(np.abs(P[0, 0] - G["C_P0"])) < 1e-12


def test_padding_path_odd_size(gold_ll):
    G = gold_ll
    pS = H.HostProblem(G["S_counts"], G["S_lags"], True)            # n = 37 -> padded to 40
    ll, _, st, _ = H.loglike_batch(pS, G["S_v"], G["S_w"], method=EIGEN)
    assert not st.any()
    np.testing.assert_allclose(ll, G["S_ll"], rtol=1e-12)


def test_nan_input_is_flagged(gold_ll):
    G = gold_ll
    pP = H.HostProblem(G["P_counts"], G["P_lags"], True)
    v = np.zeros((1, 100))
    v[0, 3] = np.nan
    for method in (EIGEN, SQUARING):
        ll, _, st, _ = H.loglike_batch(pP, v, G["P_flat_w"][None, :], method=method)
        assert np.isnan(ll[0]) and st[0] != 0


def _run_case(gold_replay, tag, method, nsteps):
    case = replay_case(gold_replay, tag)
    g = case["g"]
    prob = H.HostProblem(case["counts"], case["lagtimes"], case["pbc"], v_basis=case["v_basis"],
                         w_basis=case["w_basis"], spring_k=case["k"], string_vecs=case["svecs"])
    ch = H.HostChains(prob, 1, g("v0"), g("w0"), g("vc0"), g("wc0"), ll0=float(g("ll0")), dv=case["dv"],
                      dw=case["dw"], dt0=case["dt0"], temp=case["temp"])
    mp = K.McParams()
    mp.num_mc_update = case["nmc_update"]
    mp.dtemp = case["dtemp"]
    mp.min_lt = float(np.min(case["lagtimes"]))
    mp.move_timezero = int(case["move_timezero"])
    mp.sample_every = 100
    mp.refresh_every = 16
    mp.use_tape = 1
    mp.method = method
    dec, llt, smp = ch.run(nsteps, mp, g("tape_u"), g("tape_idx"), nsamples_cap=max(1, nsteps // 100))
    return case, g, ch, dec, llt, smp


@pytest.mark.parametrize("tag,method,nsteps", [("cosA", EIGEN, 40), ("cosA", SQUARING, 110), ("binP4", SQUARING_DFMA, 40),
                                               ("cosP4t0", EIGEN, 40), ("binCut", SQUARING, 110),
                                               ("springP4", EIGEN, 30)])
def test_replay_prefix_matches_reference(gold_replay, tag, method, nsteps):
    case, g, ch, dec, llt, smp = _run_case(gold_replay, tag, method, nsteps)
    assert np.array_equal(dec[0] >> 2, g("move")[:nsteps])
    assert np.array_equal(dec[0] & 1, g("accepted")[:nsteps])
    assert np.array_equal((dec[0] >> 1) & 1, g("accepted_t0")[:nsteps])
    np.testing.assert_allclose(ch.scal[0, K.S_LL], g("ll_after")[nsteps - 1], rtol=1e-11)
    assert ch.counters[0, K.C_STATUS] == 0 and ch.counters[0, K.C_STEP] == nsteps
    if nsteps >= 100:
        np.testing.assert_allclose(smp[0, 0, 0], g("log_ll")[1], rtol=1e-11)
        w_ref = g("widths")
        np.testing.assert_allclose(ch.scal[0, K.S_DV:K.S_DW + 1], w_ref[0, 1:], rtol=1e-13)


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10, device code vs the host mirror."""
    import ctypes as C
    from mcdiff_b200 import philox
    out = (C.c_uint * 4)()
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        H.lib().emul_philox(*ctr, *key, out)
        assert tuple(out) == want
        assert philox.philox4x32_10(*ctr, *key) == want


def test_fast_log_accuracy():
    """The branch-free logarithm of the likelihood epilogue vs libm on its whole input range."""
    import ctypes as C
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-10.5, 1.5, 200000), 1 + rng.uniform(-1e-3, 1e-3, 10000),
                        np.array([1.0, 0.5, 2.0, 1e-10, 1 - 2.0 ** -53, 1 + 2.0 ** -52])])
    y = np.zeros_like(x)
    H.lib().emul_fast_log(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_int(len(x)))
    ref = np.log(x)
    # 1 ulp of the result (libm's own error is up to 1 ulp as well)
    assert np.all(np.abs(y - ref) <= 4e-16 + 2.3e-16 * np.abs(ref))


def _blocked_problem(n, lags, pbc, seed=3):
    from oracle import mcdiff_oracle as O
    rng = np.random.default_rng(seed)
    if pbc:
        counts = O.synthetic_counts(n, lags, seed=seed, total=2.0e5, dz=1.0)[0]
    else:
        counts = rng.integers(0, 50, size=(len(lags), n, n)).astype(np.int32)
        ii, jj = np.indices((n, n))
        counts[:, np.abs(ii - jj) > 12] = 0
    return counts, rng


@pytest.mark.parametrize("n,lags,pbc", [(120, (5.0, 10.0, 15.0), True), (200, (5.0, 10.0, 20.0, 40.0), True),
                                        (136, (3.0, 7.0, 10.0), False), (106, (10.0,), True),
                                        (208, (5.0, 10.0), True), (250, (5.0, 10.0), True),
                                        (330, (4.0, 8.0, 12.0), False)])
def test_blocked_squaring_engine(n, lags, pbc):
    """104 < n <= 208: matrices in a global workspace, 2 x 2 block products through the
    shared-memory panels (fast plan, general plans with parked operands, non-periodic, a single
    lag, the largest size) against the oracle."""
    from oracle import mcdiff_oracle as O
    lags = np.array(lags)
    counts, rng = _blocked_problem(n, lags, pbc)
    dim_w = n if pbc else n - 1
    v = rng.normal(0, 0.5, (2, n))
    w = rng.normal(-1.0, 0.3, (2, dim_w))
    prob = H.HostProblem(counts, lags, pbc)
    ll, P, st, nsq = H.loglike_batch(prob, v, w, want_P=True, method=SQUARING)
    assert not st.any() and (nsq > 0).all()
    for b in range(2):
        np.testing.assert_allclose(ll[b], O.log_like_lag(v[b], w[b], lags, counts, pbc), rtol=1e-12)
        for l, lt in enumerate(lags):
            Pref = O.propagator(v[b], w[b], lt, pbc)
            assert np.max(np.abs(P[b, l] - Pref)) < 1e-13
    if n <= 208:      # (the emulated Jacobi sweeps of a 330 x 330 matrix take a minute)
        ll_e, _, st, _ = H.loglike_batch(prob, v[:1], w[:1], method=EIGEN)      # same problem, global-memory eigen path
        np.testing.assert_allclose(ll_e[0], ll[0], rtol=1e-9)   # eigen route: absolute, not componentwise, accuracy


def test_blocked_engine_mc_replay():
    """MC loop on the blocked engine (n = 120, cosine model, time-offset moves, lag set that
    needs a parked operand): decisions equal the oracle replay of the same tape."""
    from oracle import mcdiff_oracle as O
    n, lags = 120, np.array([5.0, 10.0, 20.0])
    counts, rng = _blocked_problem(n, lags, True, seed=5)
    vb, wb = O.cos_basis_center(n, 5), O.cos_basis_border(n, n, 4)
    wc0 = np.zeros(4)
    wc0[0] = -1.0
    w0 = wb @ wc0
    cfg = O.ChainConfig(counts, lags, True, v_basis=vb, w_basis=wb, dv=0.2, dw=0.2, dtimezero=0.3, temp=1.0, nmc=24,
                        num_mc_update=10.0, move_timezero=True)
    tu, ti = O.make_tape(np.random.default_rng(11), 24, cfg)
    ref = O.replay(cfg, np.zeros(n), w0, tu, ti, v_coeff0=np.zeros(5), w_coeff0=wc0)
    prob = H.HostProblem(counts, lags, True, v_basis=vb, w_basis=wb)
    ll0 = O.log_like_lag(np.zeros(n), w0, lags, counts, True)
    ch = H.HostChains(prob, 1, np.zeros(n), w0, np.zeros(5), wc0, ll0=ll0, dv=0.2, dw=0.2, dt0=0.3)
    mp = K.McParams()
    mp.num_mc_update, mp.dtemp, mp.min_lt, mp.move_timezero = 10.0, 0.0, 5.0, 1
    mp.sample_every, mp.refresh_every, mp.use_tape, mp.method = 100, 16, 1, SQUARING
    dec, llt, _ = ch.run(24, mp, tu, ti)
    assert np.array_equal(dec[0] & 1, ref["accepted"])
    assert np.array_equal((dec[0] >> 1) & 1, ref["accepted_t0"])
    assert ref["accepted"].any() and ref["accepted_t0"].any()
    np.testing.assert_allclose(ch.scal[0, K.S_LL], ref["ll"], rtol=1e-11)
    np.testing.assert_allclose(ch.scal[0, K.S_T0], ref["timezero"], rtol=1e-13)


def test_pull_general_engine(gold_pull):
    """--pull: the non-symmetric generator on the general squaring engine (full m8n8k4 products on R itself)
    against values of the unmodified reference; also sizes with padding, a single lag, a general lag plan."""
    from oracle import mcdiff_oracle as O
    G = gold_pull
    for b in range(4):
        prob = H.HostProblem(G["counts"], G["lags"], True, pull=float(G["ll/dF"][b]))
        ll, P, st, nsq = H.loglike_batch(prob, G["ll/v"][b], G["ll/w"][b], want_P=(b == 0), method=SQUARING)
        assert not st.any() and nsq[0] > 0
        np.testing.assert_allclose(ll[0], G["ll/ll"][b], rtol=1e-12)
        if b == 0:
            assert np.max(np.abs(P[0, 0] - G["P0"])) < 1e-13
            np.testing.assert_allclose(P[0, 3].sum(axis=0), 1.0, atol=1e-12)
    rng = np.random.default_rng(9)
    for n, lags, dF in [(37, (3.0, 6.0), -0.2), (60, (5.0, 10.0, 20.0), 0.1), (8, (2.0,), 0.3)]:
        lags = np.array(lags)
        counts = O.synthetic_counts(n, lags, seed=4, total=2.0e5, dz=1.0)[0]
        v, w = rng.normal(0, 0.5, n), rng.normal(-1.0, 0.3, n)
        ll, _, st, _ = H.loglike_batch(H.HostProblem(counts, lags, True, pull=dF), v, w, method=SQUARING)
        np.testing.assert_allclose(ll[0], O.log_like_lag(v, w, lags, counts, True, pull=dF), rtol=1e-12)


@pytest.mark.parametrize("tag,nsteps", [("cosPull", 60), ("binPull", 40)])
def test_pull_replay_prefix_matches_reference(gold_pull, tag, nsteps):
    case = replay_case(gold_pull, tag)
    g = case["g"]
    prob = H.HostProblem(case["counts"], case["lagtimes"], True, v_basis=case["v_basis"], w_basis=case["w_basis"],
                         pull=case["pull"])
    ch = H.HostChains(prob, 1, g("v0"), g("w0"), g("vc0"), g("wc0"), ll0=float(g("ll0")), dv=case["dv"], dw=case["dw"],
                      dt0=case["dt0"], temp=case["temp"])
    mp = K.McParams()
    mp.num_mc_update, mp.dtemp, mp.min_lt = case["nmc_update"], case["dtemp"], float(np.min(case["lagtimes"]))
    mp.move_timezero, mp.sample_every, mp.refresh_every, mp.use_tape, mp.method = int(case["move_timezero"]), 100, 16, 1, SQUARING
    dec, llt, _ = ch.run(nsteps, mp, g("tape_u"), g("tape_idx"))
    assert np.array_equal(dec[0] >> 2, g("move")[:nsteps])
    assert np.array_equal(dec[0] & 1, g("accepted")[:nsteps])
    assert np.array_equal((dec[0] >> 1) & 1, g("accepted_t0")[:nsteps])
    np.testing.assert_allclose(ch.scal[0, K.S_LL], g("ll_after")[nsteps - 1], rtol=1e-11)


def test_chain_body_with_128_threads():
    """The library compiles the same chain body for 128-thread CTAs (mcd_small.cu, profiles with np <= 56): every
    table, reduction and thread map must be free of the 384-thread assumption.  n = 48 and n = 21, cosine model with
    time-offset moves, against the oracle replay."""
    from oracle import mcdiff_oracle as O
    H.use_threads(128)
    try:
        for n, nsteps in ((48, 40), (21, 30)):
            lags = np.array([5.0, 10.0, 15.0])
            counts = O.synthetic_counts(n, lags, seed=8, total=1.0e5, dz=1.0)[0]
            vb, wb = O.cos_basis_center(n, 5), O.cos_basis_border(n, n, 4)
            wc0 = np.zeros(4)
            wc0[0] = -1.0
            w0 = wb @ wc0
            cfg = O.ChainConfig(counts, lags, True, v_basis=vb, w_basis=wb, dv=0.3, dw=0.3, dtimezero=0.3, temp=1.0,
                                nmc=nsteps, num_mc_update=10.0, move_timezero=True)
            tu, ti = O.make_tape(np.random.default_rng(n), nsteps, cfg)
            ref = O.replay(cfg, np.zeros(n), w0, tu, ti, v_coeff0=np.zeros(5), w_coeff0=wc0)
            prob = H.HostProblem(counts, lags, True, v_basis=vb, w_basis=wb)
            ll0 = O.log_like_lag(np.zeros(n), w0, lags, counts, True)
            ll, _, st, _ = H.loglike_batch(prob, np.zeros(n), w0, method=SQUARING)
            np.testing.assert_allclose(ll[0], ll0, rtol=1e-12)
            ch = H.HostChains(prob, 1, np.zeros(n), w0, np.zeros(5), wc0, ll0=ll0, dv=0.3, dw=0.3, dt0=0.3)
            mp = K.McParams()
            mp.num_mc_update, mp.dtemp, mp.min_lt, mp.move_timezero = 10.0, 0.0, 5.0, 1
            mp.sample_every, mp.refresh_every, mp.use_tape, mp.method = 100, 16, 1, SQUARING
            dec, _, _ = ch.run(nsteps, mp, tu, ti)
            assert np.array_equal(dec[0] & 1, ref["accepted"])
            assert np.array_equal((dec[0] >> 1) & 1, ref["accepted_t0"])
            np.testing.assert_allclose(ch.scal[0, K.S_LL], ref["ll"], rtol=1e-11)
            np.testing.assert_allclose(ch.scal[0, K.S_DV:K.S_DW + 1], [ref["dv"], ref["dw"]], rtol=1e-13)
    finally:
        H.use_threads(None)


# ---- time offsets on the squaring engines (plan selection, emulator mirrors the kernels) -------------
@pytest.mark.parametrize("lags", [[10.0, 20.0, 30.0, 40.0], [5.0, 10.0, 25.0]])
def test_shifted_lag_likelihood_on_squaring_engine(lags):
    """Per-profile lag times = the problem's lag times + one offset walk plan[1] on the squaring engine and agree
    with the oracle (scipy expm at the shifted lag times); anything else is flagged MCD_ST_BAD_LAGS."""
    from oracle import mcdiff_oracle as O
    n = 24
    lags = np.array(lags)
    counts = O.synthetic_counts(n, lags, seed=5, total=5.0e4, dz=1.0)[0]
    prob = H.HostProblem(counts, lags, True)
    rng = np.random.default_rng(1)
    v = rng.normal(0.0, 0.5, (3, n))
    w = -1.0 + rng.normal(0.0, 0.3, (3, n))
    shift = np.array([3.0, -1.5, 0.0])
    lt = lags[None, :] + shift[:, None]
    ll, _, st, _ = H.loglike_batch(prob, v, w, lagtimes=lt, method=SQUARING)
    assert not st.any()
    ref = np.array([O.log_like_lag(v[b], w[b], lt[b], counts, True) for b in range(3)])
    np.testing.assert_allclose(ll, ref, rtol=1e-11)
    ll_e, _, st_e, _ = H.loglike_batch(prob, v, w, lagtimes=lt, method=EIGEN)
    np.testing.assert_allclose(ll_e, ref, rtol=1e-9)
    bad = lt.copy()
    bad[1, 1] += 0.5                                   # not an offset of the problem's lag times
    ll_b, _, st_b, _ = H.loglike_batch(prob, v, w, lagtimes=bad, method=SQUARING)
    assert st_b[1] == K.ST_BAD_LAGS and np.isnan(ll_b[1]) and st_b[0] == 0 and st_b[2] == 0
    np.testing.assert_allclose(ll_b[[0, 2]], ref[[0, 2]], rtol=1e-11)


@pytest.mark.parametrize("method", [EIGEN, SQUARING])
def test_time_offset_without_time_offset_moves(method):
    """A chain that CARRIES a time offset (continued --t0 run, set_state(timezero=...)) but makes no time-offset
    moves must evaluate its trials at lag + t0: on the squaring engines that is the shifted LagPlan."""
    from oracle import mcdiff_oracle as O
    n, nsteps, tau = 24, 30, 3.0
    lags = np.array([10.0, 20.0, 30.0, 40.0])
    counts = O.synthetic_counts(n, lags, seed=7, total=5.0e4, dz=1.0)[0]
    prob = H.HostProblem(counts, lags, True)
    v0, w0 = np.zeros(n), np.full(n, -1.0)
    ll0 = O.log_like_lag(v0, w0, lags + tau, counts, True)
    ch = H.HostChains(prob, 1, v0, w0, ll0=ll0, dv=0.3, dw=0.3, timezero=tau)
    cfg = O.ChainConfig(counts, lags, True, dv=0.3, dw=0.3, temp=1.0, nmc=nsteps, num_mc_update=10.0)
    tu, ti = O.make_tape(np.random.default_rng(3), nsteps, cfg)
    mp = K.McParams()
    mp.num_mc_update, mp.min_lt, mp.move_timezero, mp.sample_every = 10.0, 10.0, 0, 100
    mp.refresh_every, mp.use_tape, mp.method = 16, 1, method
    dec, llt, _ = ch.run(nsteps, mp, tu, ti)
    ref = O.replay(cfg, v0, w0, tu, ti, timezero0=tau)
    assert np.array_equal(dec[0] & 1, ref["accepted"])
    np.testing.assert_allclose(llt[0], ref["ll_try"], rtol=1e-10)
    np.testing.assert_allclose(ch.scal[0, K.S_LL], ref["ll"], rtol=1e-11)
    assert ch.scal[0, K.S_T0] == tau and not ch.counters[0, K.C_STATUS]


# ---- packed symmetric engine (mcd_pk.cu: 192 threads, tile-packed work matrices, two chains per SM) ---------------
PACKED = 4      # emulator-only method code (the library picks this engine by itself for 57 <= np <= 104)


def test_packed_engine_likelihoods(gold_ll):
    """Likelihoods and propagators of the packed engine, compiled as the library compiles it (192 threads: products
    run in two rounds), on the reference fixtures: one lag, four lags (fast plan), non-periodic, odd size."""
    G = gold_ll
    H.use_threads(256)
    try:
        pA = H.HostProblem(G["A_counts"], G["A_lags"], True)
        ll, P, st, _ = H.loglike_batch(pA, G["K_v"][:2], G["K_w"][:2], want_P=True, method=PACKED)
        assert not st.any()
        np.testing.assert_allclose(ll, G["K_ll"][:2], rtol=1e-12)
        assert np.max(np.abs(P[1, 0] - G["K2_P"])) < 1e-13
        pP = H.HostProblem(G["P_counts"], G["P_lags"], True)
        ll, _, st, _ = H.loglike_batch(pP, G["rand_v"][:3], G["rand_wP"][:3], method=PACKED)
        np.testing.assert_allclose(ll, G["rand_ll_P"][:3], rtol=1e-12)
        pC = H.HostProblem(G["C_counts"], G["C_lags"], False)
        ll, P, st, _ = H.loglike_batch(pC, G["C_v"][:2], G["C_w"][:2], want_P=True, method=PACKED)
        np.testing.assert_allclose(ll, G["C_ll"][:2], rtol=1e-12)
        assert np.max(np.abs(P[0, 0] - G["C_P0"])) < 1e-12
    finally:
        H.use_threads(None)


@pytest.mark.parametrize("n,lags,pbc", [(61, [3.0, 6.0, 9.0, 12.0, 15.0, 18.0], True),     # long progression: scratch + copy back
                                        (72, [2.0, 3.0, 5.0, 10.0], True),                  # general plan with slots
                                        (100, [4.0, 4.0, 8.0], False),                      # duplicate lag time
                                        (97, [10.0], True)])
def test_packed_engine_plans(n, lags, pbc):
    from oracle import mcdiff_oracle as O
    lags = np.array(lags)
    counts = O.synthetic_counts(n, lags, seed=n, total=2.0e5, dz=0.7)[0]
    if not pbc:
        counts = counts.copy()
        counts[:, 0, n - 1] = 0
        counts[:, n - 1, 0] = 0
    rng = np.random.default_rng(n)
    v = rng.normal(0.0, 0.6, (2, n))
    w = -0.5 + rng.normal(0.0, 0.3, (2, n if pbc else n - 1))
    H.use_threads(256)
    try:
        prob = H.HostProblem(counts, lags, pbc)
        ll, P, st, _ = H.loglike_batch(prob, v, w, want_P=True, method=PACKED)
        assert not st.any()
        ref = np.array([O.log_like_lag(v[b], w[b], lags, counts, pbc) for b in range(2)])
        np.testing.assert_allclose(ll, ref, rtol=1e-11)
        for l, lag in enumerate(lags):
            assert np.max(np.abs(P[0, l] - O.propagator(v[0], w[0], lag, pbc))) < 1e-12
        if len(lags) > 1 and pbc:
            shift = np.array([0.7, -0.4])
            lt = lags[None, :] + shift[:, None]
            ll_s, _, st_s, _ = H.loglike_batch(prob, v, w, lagtimes=lt, method=PACKED)
            assert not st_s.any()
            ref_s = np.array([O.log_like_lag(v[b], w[b], lt[b], counts, pbc) for b in range(2)])
            np.testing.assert_allclose(ll_s, ref_s, rtol=1e-11)
    finally:
        H.use_threads(None)


def test_packed_engine_mc_replay(gold_replay):
    """The MC loop on the packed engine against a reference trajectory (cosine model on A, 110 steps) and, with
    time-offset moves, against the oracle replay at the BASELINE shape."""
    from oracle import mcdiff_oracle as O
    H.use_threads(256)
    try:
        case, g, ch, dec, llt, smp = _run_case(gold_replay, "cosA", PACKED, 110)
        assert np.array_equal(dec[0] >> 2, g("move")[:110])
        assert np.array_equal(dec[0] & 1, g("accepted")[:110])
        np.testing.assert_allclose(ch.scal[0, K.S_LL], g("ll_after")[109], rtol=1e-11)
        assert ch.counters[0, K.C_STATUS] == 0
        n, nsteps = 100, 24
        lags = np.array([10.0, 20.0, 30.0, 40.0])
        counts = O.synthetic_counts(n, lags, seed=11, total=2.0e5, dz=0.5)[0]
        vb, wb = O.cos_basis_center(n, 6), O.cos_basis_border(n, n, 4)
        wc0 = np.zeros(4)
        wc0[0] = 1.0
        w0 = wb @ wc0
        cfg = O.ChainConfig(counts, lags, True, v_basis=vb, w_basis=wb, dv=0.2, dw=0.2, dtimezero=0.3, temp=1.0,
                            nmc=nsteps, num_mc_update=8.0, move_timezero=True)
        tu, ti = O.make_tape(np.random.default_rng(2), nsteps, cfg)
        ref = O.replay(cfg, np.zeros(n), w0, tu, ti, v_coeff0=np.zeros(6), w_coeff0=wc0)
        prob = H.HostProblem(counts, lags, True, v_basis=vb, w_basis=wb)
        ll0 = O.log_like_lag(np.zeros(n), w0, lags, counts, True)
        chn = H.HostChains(prob, 1, np.zeros(n), w0, np.zeros(6), wc0, ll0=ll0, dv=0.2, dw=0.2, dt0=0.3)
        mp = K.McParams()
        mp.num_mc_update, mp.dtemp, mp.min_lt, mp.move_timezero = 8.0, 0.0, 10.0, 1
        mp.sample_every, mp.refresh_every, mp.use_tape, mp.method = 100, 16, 1, PACKED
        dec, _, _ = chn.run(nsteps, mp, tu, ti)
        assert np.array_equal(dec[0] & 1, ref["accepted"])
        assert np.array_equal((dec[0] >> 1) & 1, ref["accepted_t0"])
        np.testing.assert_allclose(chn.scal[0, K.S_LL], ref["ll"], rtol=1e-11)
    finally:
        H.use_threads(None)


# ---- tensor-only build of the full-storage engine (Chain<Exec, 4>: mcd_mc_run_sq_kernel, mcd128_mc_run_kernel) ------
TENSOR_ONLY = 5  # emulator-only method code


@pytest.mark.parametrize("threads", [None, 128])
@pytest.mark.parametrize("n,lags,pbc", [(5, [1.0, 2.0, 3.0, 5.0, 10.0], True),      # general plan with slots (fuzz case)
                                        (24, [4.0, 4.0, 8.0], False),               # duplicate lag: buffer_loglike
                                        (40, [1.0e-4, 2.0e-4], True),               # kernels that need no squaring
                                        (56, [10.0, 20.0, 30.0, 40.0], True)])
def test_tensor_only_engine_plans(threads, n, lags, pbc):
    """Every branch of evaluate_sq in the production build: parked operands, duplicate lag times and squaring-free
    kernels take their likelihood terms from buffer_loglike, which walks the 4x4 tile table -- the table has to exist
    in this build too (it did not: illegal addresses on the GPU, found by the randomised sweep)."""
    from oracle import mcdiff_oracle as O
    lags = np.array(lags)
    counts = O.synthetic_counts(n, np.maximum(lags, 1.0), seed=n, total=1.0e5, dz=0.7)[0]
    if not pbc:
        counts = counts.copy()
        counts[:, 0, n - 1] = 0
        counts[:, n - 1, 0] = 0
    rng = np.random.default_rng(n)
    v = rng.normal(0.0, 0.6, (2, n))
    w = -0.5 + rng.normal(0.0, 0.3, (2, n if pbc else n - 1))
    H.use_threads(threads)
    try:
        prob = H.HostProblem(counts, lags, pbc)
        ll, P, st, _ = H.loglike_batch(prob, v, w, want_P=True, method=TENSOR_ONLY)
        assert not st.any()
        ref = np.array([O.log_like_lag(v[b], w[b], lags, counts, pbc) for b in range(2)])
        np.testing.assert_allclose(ll, ref, rtol=1e-11)
        for l, lag in enumerate(lags):
            assert np.max(np.abs(P[0, l] - O.propagator(v[0], w[0], lag, pbc))) < 1e-12
        if pbc and len(lags) > 2:
            lt = lags[None, :] + np.array([0.3, -0.2])[:, None]                 # time-offset trial: plan[1]
            ll_s, _, st_s, _ = H.loglike_batch(prob, v, w, lagtimes=lt, method=TENSOR_ONLY)
            assert not st_s.any()
            ref_s = np.array([O.log_like_lag(v[b], w[b], lt[b], counts, pbc) for b in range(2)])
            np.testing.assert_allclose(ll_s, ref_s, rtol=1e-11)
    finally:
        H.use_threads(None)


def test_tensor_only_engine_mc_replay():
    """The MC loop of the production build on the host: general lag plan + time-offset moves at n = 5 (the failing
    configuration of the randomised GPU sweep) against the oracle replay, 128-thread build."""
    from mcdiff_b200 import _capi as K, philox
    from oracle import mcdiff_oracle as O
    n, lags, seed, nsteps = 5, np.array([1.0, 2.0, 3.0, 5.0, 10.0]), 736475806, 25
    rng = np.random.default_rng(6)
    counts = (rng.poisson(3.0, size=(len(lags), n, n)) * 30).astype(np.int64)
    v0, w0 = np.zeros(n), np.full(n, -1.0)
    H.use_threads(128)
    try:
        prob = H.HostProblem(counts, lags, True)
        ll0 = O.log_like_lag(v0, w0, lags, counts, True)
        ch = H.HostChains(prob, 2, v0, w0, ll0=ll0, dv=0.3, dw=0.3, dt0=0.3)
        mp = K.McParams()
        mp.num_mc_update, mp.min_lt, mp.move_timezero = 10.0, 1.0, 1
        mp.sample_every, mp.refresh_every, mp.seed, mp.method = 100, 512, seed, TENSOR_ONLY
        dec, _ = ch.run(nsteps, mp)[:2]
        cfg = O.ChainConfig(counts, lags, True, dv=0.3, dw=0.3, dtimezero=0.3, temp=1.0, nmc=nsteps, num_mc_update=10.0,
                            move_timezero=True)
        for c in range(2):
            tu, ti = philox.chain_tape(seed, c, 0, nsteps, n, n, 0, 0)
            out = O.replay(cfg, v0, w0, tu, ti)
            assert np.array_equal(dec[c] & 1, out["accepted"]) and np.array_equal((dec[c] >> 1) & 1, out["accepted_t0"])
            assert abs(ch.scal[c][0] - out["ll"]) <= 1e-9 * abs(out["ll"])
    finally:
        H.use_threads(None)


@pytest.mark.parametrize("n,pbc,lag", [(31, True, 11.0), (64, False, 7.0), (100, True, 53.0)])
def test_rough_profiles_on_the_production_build(n, pbc, lag):
    """Rough profiles (v ~ N(0, 1.5), w ~ N(-1, 0.8): rates spread over three decades, 4..9 squarings, arguments in the
    stretched Taylor range (2, 2.83]) on the production build of the squaring engine: log L to 1e-12 relative and
    every propagator entry >= 1e-10 to 1e-9 relative (measured 8e-15 and 3e-13, scipy's own error included)."""
    from oracle import mcdiff_oracle as O
    rng = np.random.default_rng(n)
    lags = np.array([lag, 2 * lag])
    counts = (rng.poisson(2.0, size=(2, n, n)) * 5).astype(np.int64)
    v = rng.normal(0.0, 1.5, (3, n))
    w = rng.normal(-1.0, 0.8, (3, n if pbc else n - 1))
    prob = H.HostProblem(counts, lags, pbc)
    ll, P, st, _ = H.loglike_batch(prob, v, w, want_P=True, method=TENSOR_ONLY)
    assert not st.any()
    ref = np.array([O.log_like_lag(v[b], w[b], lags, counts, pbc) for b in range(3)])
    np.testing.assert_allclose(ll, ref, rtol=1e-12)
    for l in range(2):
        Pref = O.propagator(v[0], w[0], lags[l], pbc)
        big = Pref >= 1e-10
        np.testing.assert_allclose(P[0, l][big], Pref[big], rtol=1e-9)
        assert np.max(np.abs(P[0, l][~big] - Pref[~big]), initial=0.0) < 1e-14


@pytest.mark.parametrize("L", [4, 5, 6, 7, 8])
def test_blocked_engine_multiples_of_the_first_lag(L):
    """Multiples t0 .. L t0 on the blocked engine (the fast plan: every lag product takes G(t0) as its operand), every
    lag's likelihood terms and propagator against the oracle -- n = 110 (2 x 2 blocks of 56), non-periodic for odd L."""
    from oracle import mcdiff_oracle as O
    n, pbc = 110, (L % 2 == 0)
    lags = 5.0 * np.arange(1, L + 1)
    counts = O.synthetic_counts(n, lags, seed=L, total=2.0e5, dz=0.6)[0]
    if not pbc:
        counts = counts.copy()
        counts[:, 0, n - 1] = 0
        counts[:, n - 1, 0] = 0
    rng = np.random.default_rng(L)
    v = rng.normal(0.0, 0.5, (2, n))
    w = -0.8 + rng.normal(0.0, 0.3, (2, n if pbc else n - 1))
    prob = H.HostProblem(counts, lags, pbc)
    ll, P, st, _ = H.loglike_batch(prob, v, w, want_P=True, method=SQUARING)
    assert not st.any()
    ref = np.array([O.log_like_lag(v[b], w[b], lags, counts, pbc) for b in range(2)])
    np.testing.assert_allclose(ll, ref, rtol=1e-11)
    for l, lag in enumerate(lags):
        assert np.max(np.abs(P[1, l] - O.propagator(v[1], w[1], lag, pbc))) < 1e-12

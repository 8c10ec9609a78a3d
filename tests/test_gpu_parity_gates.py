"""BASELINE.md 4.5 parity gates, as driver-run GPU tests (through the C ABI):

  G1  1 024 random profiles on the HexdWat fixture (A: K1/K2's transition matrix, M: K3/K4's) against the oracle,
      log L relative error <= 1e-9 on every engine;
  G2  propagator entries: relative error <= 1e-9 on EVERY entry >= the clip 1e-10 against a componentwise-accurate
      long-double propagator (scipy's expm is only ~1e-16 ABSOLUTE, so it cannot judge small entries), and
      |dP| <= 1e-15 + 1e-9 P against the reference's scipy.linalg.expm golden;
  G3  >= 10 000 accept/reject decisions (8 chains x 700 steps x {profile move, time-offset move}) on device Philox
      streams at the BASELINE shape, identical to the oracle replay;
  G4  the MC loop at config 3's exact shape -- 200 bins x 8 lags, CosinusModel ncosF 12 / ncosD 8 -- on the blocked
      engine, decisions identical to the oracle replay.
"""
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9
METHODS = {"auto": 0, "eigen": 1, "squaring": 2, "squaring_dfma": 3}


# ---------------------------------------------------------------------------------------------------------------
# G1
# ---------------------------------------------------------------------------------------------------------------
def _oracle_ll(args):
    from oracle import mcdiff_oracle as O
    v, w, lags, counts = args
    return [O.log_like_lag(v[b], w[b], lags, counts, True) for b in range(len(v))]


def _oracle_batch(v, w, lags, counts):
    nproc = min(os.cpu_count() or 1, 16)
    parts = np.array_split(np.arange(len(v)), nproc)
    with ProcessPoolExecutor(max_workers=nproc) as ex:
        out = ex.map(_oracle_ll, [(v[p], w[p], lags, counts) for p in parts if len(p)])
    return np.concatenate([np.asarray(o) for o in out])


@pytest.mark.parametrize("fixture,nprof", [("A", 1024), ("M", 256)])
def test_1024_random_profiles_on_hexdwat(engine, gold_ll, fixture, nprof):
    from mcdiff_b200.engine import DeviceProblem
    G = gold_ll
    counts, lags = G[fixture + "_counts"], G[fixture + "_lags"]
    n = counts.shape[-1]
    rng = np.random.default_rng(1024 if fixture == "A" else 256)
    kw = G["K_w"][0 if fixture == "A" else 2]
    # the spread of the reference's own profiles: F within a few kT, ln D within +-1 of the flat start; a quarter of
    # the profiles are rough (sigma 1 / 0.5), a quarter smooth cosines, the rest in between
    v = rng.normal(0.0, 1.0, (nprof, n)) * rng.choice([0.1, 0.5, 1.0, 2.0], (nprof, 1))
    w = kw[None, :] + rng.normal(0.0, 1.0, (nprof, n)) * rng.choice([0.05, 0.3, 0.5, 1.0], (nprof, 1))
    x = (np.arange(n) + 0.5) / n
    smooth = np.arange(nprof) % 4 == 0
    v[smooth] = 3.0 * np.cos(2 * np.pi * x)[None, :] * rng.normal(0.0, 1.0, (int(smooth.sum()), 1))
    ref = _oracle_batch(v, w, lags, counts)
    assert np.all(np.isfinite(ref))
    prob = DeviceProblem(engine, counts, lags, True)
    worst = {}
    for name in ("auto", "squaring", "squaring_dfma", "eigen"):
        ll, _, st = prob.loglike(v, w, method=METHODS[name])
        assert not st.any(), name
        worst[name] = float(np.max(np.abs(ll - ref) / np.abs(ref)))
        # The squaring engines sum non-negative terms only (componentwise accurate, ~1e-15 here).  The eigen engine
        # (Jacobi, the fallback for lag sets without a LagPlan) forms P = V e^{t Lambda} V^T with cancellation: it is
        # accurate NORM-wise like any eigendecomposition, so an entry near the clip 1e-10 carries a relative error
        # ~1e-16 / 1e-10 and rough profiles (many such entries with counts) reach ~2e-8 in log L.
        np.testing.assert_allclose(ll, ref, rtol=RTOL if name != "eigen" else 1e-7, err_msg=name)
    assert worst["auto"] <= 1e-12 and worst["squaring"] <= 1e-12      # what the default engine actually achieves


# ---------------------------------------------------------------------------------------------------------------
# G2
# ---------------------------------------------------------------------------------------------------------------
def _expm_componentwise(R, t):
    """exp(tR) of a rate matrix (non-negative off-diagonal) with COMPONENTWISE relative accuracy: uniformisation
    exp(tR) = e^{-t sigma} exp(t (R + sigma I)) makes every term of the scaled Taylor series and of the squarings
    non-negative, evaluated in long double (64-bit mantissa)."""
    ld = np.longdouble
    R = R.astype(ld)
    n = R.shape[0]
    sigma = np.max(-np.diag(R))
    A = R + sigma * np.eye(n, dtype=ld)
    s = max(0, int(np.ceil(np.log2(float(t * np.max(A.sum(axis=0)))))) + 1)
    H = A * (ld(t) / ld(2 ** s))
    T = np.eye(n, dtype=ld)
    term = np.eye(n, dtype=ld)
    for k in range(1, 40):
        term = term.dot(H) / ld(k)
        T = T + term
    for _ in range(s):
        T = T.dot(T)
    return T * np.exp(-ld(t) * sigma)


@pytest.mark.parametrize("method", ["squaring", "squaring_dfma", "eigen"])
def test_propagator_every_entry_above_the_clip(engine, gold_ll, method):
    from mcdiff_b200.engine import DeviceProblem
    from oracle import mcdiff_oracle as O
    G = gold_ll
    pA = DeviceProblem(engine, G["A_counts"], G["A_lags"], True)
    lag = float(G["A_lags"][0])
    for v, w, Pgold in ((G["K_v"][1], G["K_w"][1], G["K2_P"]), (G["rand_v"][0], G["rand_w"][0], G["rand_P0"]),
                        (G["rand_v"][5], G["rand_w"][5], None)):
        _, P, st = pA.loglike(v, w, want_P=True, method=METHODS[method])
        P = P[0, 0]
        assert not st.any()
        truth = _expm_componentwise(O.rate_matrix(v, w, True), lag).astype(np.float64)
        if Pgold is not None:      # the reference's own scipy.linalg.expm: absolute accuracy ~1e-16
            assert np.all(np.abs(P - Pgold) <= 1e-15 + RTOL * Pgold)
            assert np.all(np.abs(truth - Pgold) <= 1e-15 + RTOL * Pgold)       # and the yardstick agrees with it
        sel = truth >= 1e-10
        assert sel.sum() > 1000
        rel = np.abs(P - truth)[sel] / truth[sel]
        # squaring engines sum non-negative terms only: componentwise accurate.  The eigen engine is accurate
        # norm-wise (V e^{t Lambda} V^T cancels), so it gets the absolute floor the reference's expm has.
        if method == "eigen":
            assert np.all(np.abs(P - truth) <= 1e-13 + RTOL * truth)
        else:
            assert rel.max() <= RTOL, rel.max()
            assert rel.max() <= 1e-12                                           # what the kernel actually achieves


# ---------------------------------------------------------------------------------------------------------------
# G3 / G4
# ---------------------------------------------------------------------------------------------------------------
def _replay_chain(args):
    from mcdiff_b200 import philox
    from oracle import mcdiff_oracle as O
    (counts, lags, ncf, ncd, wc0, dv, dw, dt0, nsteps, nupd, move_t0, seed, gchain) = args
    n = counts.shape[-1]
    vb, wb = O.cos_basis_center(n, ncf), O.cos_basis_border(n, n, ncd)
    cfg = O.ChainConfig(counts, lags, True, v_basis=vb, w_basis=wb, dv=dv, dw=dw, dtimezero=dt0, temp=1.0,
                        nmc=nsteps, num_mc_update=nupd, move_timezero=move_t0)
    tu, ti = philox.chain_tape(seed, gchain, 0, nsteps, n, n, ncf, ncd)
    # one BLAS thread per worker: several replays run side by side, and 16 OpenBLAS threads each on a 16-core host
    # make a 200 x 200 expm ten times slower than a single thread does
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=1):
        out = O.replay(cfg, np.zeros(n), wb @ wc0, tu, ti, v_coeff0=np.zeros(ncf), w_coeff0=wc0)
    return out["accepted"], out["accepted_t0"], out["move"], out["ll"], out["timezero"], out["dv"], out["dw"]


def _device_vs_replay(engine, counts, lags, ncf, ncd, wc00, dv, dw, dt0, nch, nsteps, nupd, move_t0, seed, off,
                      chunks, method, replay=True):
    from mcdiff_b200 import _capi as K
    from mcdiff_b200.engine import ChainBatch, DeviceProblem
    from oracle import mcdiff_oracle as O
    n = counts.shape[-1]
    vb, wb = O.cos_basis_center(n, ncf), O.cos_basis_border(n, n, ncd)
    wc0 = np.zeros(ncd)
    wc0[0] = wc00
    prob = DeviceProblem(engine, counts, lags, True, v_basis=vb, w_basis=wb)
    ch = ChainBatch(prob, nch)
    ch.set_state(np.zeros(n), wb @ wc0, np.zeros(ncf), wc0, dv=dv, dw=dw, dtimezero=dt0)
    ch.init_log_like()
    mp = K.McParams()
    mp.num_mc_update = nupd
    mp.dtemp = 0.0
    mp.min_lt = float(np.min(lags))
    mp.move_timezero = int(move_t0)
    mp.sample_every = 100
    mp.refresh_every = 512
    mp.use_tape = 0
    mp.seed = seed
    mp.chain_offset = off
    mp.method = method
    decs = []
    for nst in chunks:                               # several launches: the Philox counter carries over
        d, _, _ = ch.run(nst, mp, record=True)
        decs.append(d.cpu().numpy())
    dec = np.concatenate(decs, axis=1)
    st = ch.host_state()
    assert not st["counters"][:, 6].any()
    if not replay:                                   # device run only (compared with another device run by the caller)
        return st["scal"].copy(), dec, prob
    jobs = [(counts, lags, ncf, ncd, wc0, dv, dw, dt0, nsteps, nupd, move_t0, seed, off + c) for c in range(nch)]
    with ProcessPoolExecutor(max_workers=min(nch, os.cpu_count() or 1)) as ex:
        refs = list(ex.map(_replay_chain, jobs))
    ndec = 0
    for c, (acc, acc_t0, move, ll, t0, rdv, rdw) in enumerate(refs):
        assert np.array_equal(dec[c] >> 2, move), "chain %d" % c
        assert np.array_equal(dec[c] & 1, acc), "chain %d" % c
        assert np.array_equal((dec[c] >> 1) & 1, acc_t0), "chain %d" % c
        np.testing.assert_allclose(st["scal"][c][0], ll, rtol=RTOL)
        np.testing.assert_allclose(st["scal"][c][1], t0, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(st["scal"][c][2:4], [rdv, rdw], rtol=1e-12)
        ndec += len(acc) * (2 if move_t0 else 1)
    return ndec, dec, prob


def test_ten_thousand_decisions_equal_the_oracle_replay(engine):
    """BASELINE shape (100 bins x 4 lags, Fourier 10/6), time-offset moves on, width adaptation every 100 steps."""
    from oracle import mcdiff_oracle as O
    counts, lags, edges, F, w_true = O.synthetic_counts()
    ndec, dec, prob = _device_vs_replay(engine, counts, lags, 10, 6, float(np.mean(w_true)), 0.05, 0.05, 0.2,
                                        nch=16, nsteps=1000, nupd=100.0, move_t0=True, seed=4242, off=0,
                                        chunks=(200, 300, 500), method=0)
    assert prob.squaring_ok and prob.in_smem
    assert ndec >= 32000
    acc = int((dec & 1).sum())
    assert 0.05 * dec.size < acc < 0.95 * dec.size                       # both outcomes are exercised
    assert int(((dec >> 1) & 1).sum()) > 0


def test_config3_shape_mc_replay(engine, monkeypatch):
    """200 bins x 8 lags {5..40}, CosinusModel ncosF 12 / ncosD 8 (BASELINE.json config 3) on the blocked engine: one SM
    per chain against the oracle replay, and the default launch -- which puts a batch smaller than a wave on two-CTA
    clusters -- against that run (same decisions, same final state)."""
    from oracle import mcdiff_oracle as O
    lags = np.array([5.0, 10.0, 15.0, 20.0, 25.0, 30.0, 35.0, 40.0])
    counts = O.synthetic_counts(200, lags, seed=3, total=1.0e6, dz=0.5)[0]
    args = (engine, counts, lags, 12, 8, -np.log(0.25), 0.05, 0.05, 0.1)
    kw = dict(nch=6, nsteps=80, nupd=12.0, move_t0=False, seed=33, off=7, chunks=(50, 30), method=0)
    monkeypatch.setenv("MCDIFF_B200_CLUSTER", "0")
    ndec, dec, prob = _device_vs_replay(*args, **kw)
    assert prob.squaring_ok and not prob.in_smem                          # the blocked engine ran it
    assert ndec == 6 * 80
    assert 0 < int((dec & 1).sum()) < dec.size
    scal0, dec0, _ = _device_vs_replay(*args, replay=False, **kw)
    monkeypatch.delenv("MCDIFF_B200_CLUSTER", raising=False)
    scal1, dec1, _ = _device_vs_replay(*args, replay=False, **kw)
    assert np.array_equal(dec0, dec) and np.array_equal(dec1, dec)
    np.testing.assert_allclose(scal1, scal0, rtol=1e-12)


def test_packed_engine_two_chains_per_sm(engine, monkeypatch):
    """The opt-in packed symmetric engine (mcd_pk.cu, MCDIFF_B200_PACKED=1: tile-packed work matrices, 256 threads, two
    chains per SM): decisions identical to the oracle replay at the BASELINE shape with time-offset moves (general
    plan: slots + scratch), and identical to the default full-storage kernel on a longer run."""
    from mcdiff_b200 import _capi as K
    from mcdiff_b200.engine import ChainBatch
    from oracle import mcdiff_oracle as O
    counts, lags, edges, F, w_true = O.synthetic_counts()
    monkeypatch.setenv("MCDIFF_B200_PACKED", "1")
    ndec, dec, prob = _device_vs_replay(engine, counts, lags, 10, 6, float(np.mean(w_true)), 0.05, 0.05, 0.2,
                                        nch=3, nsteps=60, nupd=20.0, move_t0=True, seed=77, off=5,
                                        chunks=(25, 35), method=0)
    assert ndec == 3 * 60 * 2 and prob.in_smem

    def run(packed):
        monkeypatch.setenv("MCDIFF_B200_PACKED", "1" if packed else "0")
        n = counts.shape[-1]
        ch = ChainBatch(prob, 300)                                  # more chains than SMs: two resident per SM
        wc0 = np.zeros(6)
        wc0[0] = float(np.mean(w_true))
        wb = O.cos_basis_border(n, n, 6)
        ch.set_state(np.zeros(n), wb @ wc0, np.zeros(10), wc0, dv=0.05, dw=0.05)
        ch.init_log_like()
        mp = K.McParams()
        mp.num_mc_update, mp.min_lt, mp.sample_every, mp.refresh_every, mp.seed = 100.0, 10.0, 100, 512, 5
        d, _, _ = ch.run(150, mp, record=True)
        return d.cpu().numpy(), ch.host_state()

    d1, s1 = run(True)
    d0, s0 = run(False)
    assert np.array_equal(d1, d0)
    np.testing.assert_allclose(s1["scal"][:, 0], s0["scal"][:, 0], rtol=1e-12)
    assert not s1["counters"][:, 6].any()


def test_blocked_engine_on_a_two_cta_cluster(engine, monkeypatch):
    """The opt-in cluster kernel of the blocked engine (MCDIFF_B200_CLUSTER=1: two SMs per chain, the CTAs split the
    blocks of every product and the Taylor columns, partial sums exchanged through the workspace) makes the decisions
    of the default one-SM kernel, step for step, at config 3's shape with time-offset moves (general plan)."""
    from mcdiff_b200 import _capi as K
    from mcdiff_b200.engine import ChainBatch, DeviceProblem
    from oracle import mcdiff_oracle as O
    lags = np.array([5.0, 10.0, 15.0, 20.0, 25.0, 30.0, 35.0, 40.0])
    counts = O.synthetic_counts(200, lags, seed=3, total=1.0e6, dz=0.5)[0]
    n = 200
    vb, wb = O.cos_basis_center(n, 12), O.cos_basis_border(n, n, 8)
    prob = DeviceProblem(engine, counts, lags, True, v_basis=vb, w_basis=wb)
    assert prob.squaring_ok and not prob.in_smem
    wc0 = np.zeros(8)
    wc0[0] = -np.log(0.25)

    def run(cluster, move_t0):
        monkeypatch.setenv("MCDIFF_B200_CLUSTER", "1" if cluster else "0")
        ch = ChainBatch(prob, 5)
        ch.set_state(np.zeros(n), wb @ wc0, np.zeros(12), wc0, dv=0.1, dw=0.1, dtimezero=0.2)
        ch.init_log_like()
        mp = K.McParams()
        mp.num_mc_update, mp.min_lt, mp.sample_every, mp.refresh_every, mp.seed = 10.0, 5.0, 100, 512, 3
        mp.move_timezero = int(move_t0)
        d, ll, _ = ch.run(24, mp, record=True)
        return d.cpu().numpy(), ll.cpu().numpy(), ch.host_state()

    for move_t0 in (False, True):
        d1, l1, s1 = run(True, move_t0)
        d0, l0, s0 = run(False, move_t0)
        assert np.array_equal(d1, d0)
        np.testing.assert_allclose(l1, l0, rtol=1e-12)
        np.testing.assert_allclose(s1["scal"], s0["scal"], rtol=1e-12)
        assert not s1["counters"][:, 6].any() and 0 < int((d1 & 1).sum()) < d1.size


def test_blocked_engine_tail_wave_on_clusters(engine, monkeypatch):
    """Default launch of the blocked engine with more chains than SMs: the full waves run one SM per chain, the chains
    of the last, partly filled wave on two-CTA clusters behind them (mcd_mc_run).  Every chain must come out exactly as
    from a launch without the split (MCDIFF_B200_CLUSTER=0)."""
    import torch
    from mcdiff_b200 import _capi as K
    from mcdiff_b200.engine import ChainBatch, DeviceProblem
    from oracle import mcdiff_oracle as O
    lags = np.array([5.0, 10.0, 15.0, 20.0, 25.0, 30.0, 35.0, 40.0])
    counts = O.synthetic_counts(200, lags, seed=3, total=1.0e6, dz=0.5)[0]
    n = 200
    vb, wb = O.cos_basis_center(n, 12), O.cos_basis_border(n, n, 8)
    prob = DeviceProblem(engine, counts, lags, True, v_basis=vb, w_basis=wb)
    wc0 = np.zeros(8)
    wc0[0] = -np.log(0.25)
    nch = torch.cuda.get_device_properties(0).multi_processor_count + 7

    def run(split):
        if split:
            monkeypatch.delenv("MCDIFF_B200_CLUSTER", raising=False)
        else:
            monkeypatch.setenv("MCDIFF_B200_CLUSTER", "0")
        ch = ChainBatch(prob, nch)
        ch.set_state(np.zeros(n), wb @ wc0, np.zeros(12), wc0, dv=0.1, dw=0.1)
        ch.init_log_like()
        mp = K.McParams()
        mp.num_mc_update, mp.min_lt, mp.sample_every, mp.refresh_every, mp.seed = 5.0, 5.0, 100, 512, 11
        launches = engine.launch_count()
        d, ll, _ = ch.run(10, mp, record=True)
        return d.cpu().numpy(), ll.cpu().numpy(), ch.host_state(), engine.launch_count() - launches

    d1, l1, s1, k1 = run(True)
    d0, l0, s0, k0 = run(False)
    assert (k1, k0) == (2, 1)                                             # the split really launched two kernels
    assert np.array_equal(d1, d0)
    np.testing.assert_allclose(l1, l0, rtol=1e-12)
    np.testing.assert_allclose(s1["scal"], s0["scal"], rtol=1e-12)
    for key in ("v", "w"):
        np.testing.assert_allclose(s1[key], s0[key], rtol=0, atol=0)
    assert not s1["counters"][:, 6].any() and 0 < int((d1 & 1).sum()) < d1.size

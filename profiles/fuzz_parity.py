"""Randomised parity sweep (GPU vs oracle): random sizes, lag sets, boundary conditions, models, spring prior, time-offset
moves, --pull.  For every configuration: batched log L and propagators against the oracle, then a short device-Philox
MC run of a few chains against the oracle replay of the same tapes (decisions must be identical).
usage: fuzz_parity.py [nconfigs] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcdiff_b200 import _capi as K, philox
from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch
from oracle import mcdiff_oracle as O



def run(ncfg=40, seed=1234, eng=None, verbose=True):
    """Returns the number of failing configurations."""
    rng = np.random.default_rng(seed)
    eng = eng or Engine(0)
    _print = print if verbose else (lambda *a, **k: None)
    LAGSETS = [(10.0,), (5.0, 10.0, 15.0, 20.0), (10.0, 20.0, 30.0, 40.0), (2.0, 4.0, 10.0), (1.0, 2.0, 3.0, 5.0, 10.0),
               (4.0, 4.0, 8.0), (3.0, 7.0, 12.0), (20.0, 10.0, 30.0)]
    bad = 0
    t_start = time.time()
    for it in range(ncfg):
        n = int(rng.choice([5, 12, 21, 30, 37, 48, 56, 57, 64, 80, 97, 100, 104, 110, 128, 150, 200, 230, 320]))
        lags = np.array(LAGSETS[rng.integers(len(LAGSETS))])
        pull = float(rng.normal(0, 0.2)) if (rng.random() < 0.25 and n <= 104 and n >= 8) else None
        pbc = True if pull is not None else bool(rng.random() < 0.7)
        dim_w = n if pbc else n - 1
        use_cos = rng.random() < 0.6
        ncF, ncD = (int(rng.integers(2, 8)), int(rng.integers(1, 6))) if use_cos else (0, 0)
        k = float(rng.uniform(1.0, 5.0)) if (not use_cos and pull is None and rng.random() < 0.3) else -1.0
        t0move = bool(rng.random() < 0.4)
        counts = rng.poisson(3.0, size=(len(lags), n, n)).astype(np.int32)
        ii, jj = np.indices((n, n))
        dist = np.abs(ii - jj); dist = np.minimum(dist, n - dist) if pbc else dist
        counts = (counts * np.exp(-0.5 * (dist / max(2.0, n / 12.0)) ** 2) * 30).astype(np.int32)
        vb = O.cos_basis_center(n, ncF) if ncF else None
        wb = O.cos_basis_border(dim_w, n, ncD) if ncD else None
        svecs = O.string_vecs(dim_w, pbc) if k > 0 else None
        tag = "n=%d L=%s pbc=%d cos=%d/%d k=%.1f t0=%d pull=%s" % (n, list(lags), pbc, ncF, ncD, k, t0move, pull)
        try:
            prob = DeviceProblem(eng, counts, lags, pbc, v_basis=vb, w_basis=wb, spring_k=k, string_vecs=svecs, pull=pull)
        except K.McdError as e:
            _print("SKIP (unsupported) %s: %s" % (tag, e)); continue
        # ---- batched likelihood ----
        v = rng.normal(0, 0.5, (2, n)); w = rng.normal(-1.0, 0.3, (2, dim_w))
        ll, P, st = prob.loglike(v, w, want_P=True)
        ref = np.array([O.log_like_lag(v[b], w[b], lags, counts, pbc, pull=pull) for b in range(2)])
        relll = np.max(np.abs(ll - ref) / np.abs(ref))
        Pref = O.propagator(v[0], w[0], lags[-1], pbc, pull=pull)
        errP = np.max(np.abs(P[0, -1] - Pref))
        ok = (not st.any()) and relll < 1e-9 and errP < 1e-11
        # ---- MC run vs replay ----
        nch, nsteps, seed = 3, 25, int(rng.integers(1, 1 << 30))
        if ncF:
            vc0 = np.zeros(ncF); wc0 = np.zeros(ncD); wc0[0] = -1.0
            v0, w0 = vb @ vc0, wb @ wc0
        else:
            vc0 = wc0 = None
            v0, w0 = np.zeros(n), np.full(dim_w, -1.0)
        ch = ChainBatch(prob, nch)
        ch.set_state(v0, w0, vc0, wc0, dv=0.3, dw=0.3, dtimezero=0.3)
        ch.init_log_like()
        mp = K.McParams(); mp.num_mc_update = 10.0; mp.min_lt = float(np.min(lags)); mp.move_timezero = int(t0move)
        mp.sample_every = 100; mp.refresh_every = 512; mp.seed = seed
        dec, _, _ = ch.run(nsteps, mp, record=True)
        dec = dec.cpu().numpy()
        sth = ch.host_state()
        cfg = O.ChainConfig(counts, lags, pbc, v_basis=vb, w_basis=wb, dv=0.3, dw=0.3, dtimezero=0.3, temp=1.0, nmc=nsteps,
                            num_mc_update=10.0, move_timezero=t0move, k=k, pull=pull)
        for c in range(nch):
            tu, ti = philox.chain_tape(seed, c, 0, nsteps, n, dim_w, ncF, ncD)
            out = O.replay(cfg, v0, w0, tu, ti, v_coeff0=vc0, w_coeff0=wc0)
            same = np.array_equal(dec[c] & 1, out["accepted"]) and np.array_equal((dec[c] >> 1) & 1, out["accepted_t0"])
            llok = abs(sth["scal"][c][0] - out["ll"]) <= 1e-9 * abs(out["ll"])
            if not (same and llok):
                ok = False
                margins = np.abs(out["margin"]) if "margin" in out else None
                _print("   chain %d: decisions equal %s, final ll %.12g vs %.12g, min |margin| %s" % (
                    c, same, sth["scal"][c][0], out["ll"], None if margins is None else np.nanmin(margins)))
        _print("%s %s  rel ll %.1e  max|dP| %.1e  engine: in_smem %d squaring_ok %d" % ("ok  " if ok else "FAIL", tag, relll, errP,
                                                                                   prob.in_smem, prob.squaring_ok))
        bad += (not ok)
    _print("%d configurations, %d failures, %.0f s" % (ncfg, bad, time.time() - t_start))
    return bad


if __name__ == "__main__":
    failures = run(int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 1234)
    sys.exit(1 if failures else 0)

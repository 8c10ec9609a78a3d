"""Randomised radial-likelihood sweep (GPU vs oracle): random n, dim_rad, lmax, lag sets, boundary condition."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcdiff_b200.engine import Engine, rad_loglike
from oracle import mcdiff_oracle as O

ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 99)
eng = Engine(0)
LAGSETS = [(5.0,), (4.0, 8.0), (5.0, 10.0, 15.0), (2.0, 4.0, 10.0), (3.0, 3.0, 9.0), (12.0, 6.0),
           tuple(np.cumsum([1.0, 1.1, 1.23, 1.37, 1.51, 1.69, 1.83, 1.97, 2.11]))]
bad = 0
t0 = time.time()
for it in range(ncfg):
    n = int(rng.choice([6, 10, 16, 24, 30, 41, 50, 56, 64, 80, 100]))
    dim_rad = int(rng.integers(3, 70))
    lmax = int(rng.integers(2, 40))
    lags = np.array(LAGSETS[rng.integers(len(LAGSETS))])
    pbc = bool(rng.random() < 0.7)
    dim_w = n if pbc else n - 1
    i = np.arange(n)
    v = rng.uniform(0.3, 1.5) * np.cos(2 * np.pi * (i + 0.5) / n)
    w = np.log(0.5 + 0.2 * np.cos(2 * np.pi * (np.arange(dim_w) + 1.0) / n))
    zeros, table = O.bessel_tables(lmax, dim_rad)
    rate = O.rate_matrix(v, w, pbc)
    wrad0 = np.full(n, np.log(rng.uniform(0.1, 0.6)))
    counts = np.zeros((len(lags), dim_rad, n, n))
    for l, lt in enumerate(lags):
        counts[l] = rng.poisson(100.0 * np.clip(O.radial_propagator(rate, wrad0, lt, zeros, table), 0.0, None))
    wrad = wrad0[None, :] + rng.normal(0, 0.3, (2, n))
    ll, P, st = rad_loglike(eng, v, w, pbc, lags, counts, table, zeros, wrad, want_P=True)
    ref = np.array([O.rad_log_like_lag(rate, wrad[b], lags, counts, zeros, table) for b in range(2)])
    rel = np.max(np.abs(ll - ref) / np.abs(ref))
    Pref = O.radial_propagator(rate, wrad[1], lags[0], zeros, table)
    errP = np.max(np.abs(P[1, 0] - Pref)) / np.max(np.abs(Pref))
    ok = (not st.any()) and rel < 1e-9 and errP < 1e-11
    if not ok or it < 5:
        print("%s n=%d dim_rad=%d lmax=%d L=%d pbc=%d  rel ll %.1e  rel dP %.1e" % ("ok  " if ok else "FAIL", n, dim_rad, lmax, len(lags), pbc, rel, errP))
    bad += (not ok)
print("%d radial configurations, %d failures, %.0f s" % (ncfg, bad, time.time() - t0))

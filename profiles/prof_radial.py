"""Config-5 shape (radial/2-D model): time per MC step of the radial loop, with the oracle port timed beside it.
usage: prof_radial.py [nchains] [nsteps] [n] [dim_rad] [lmax] [nlag]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcdiff_b200.engine import Engine, RadialChains
from oracle import mcdiff_oracle as O

arg = lambda i, d: int(sys.argv[i]) if len(sys.argv) > i else d  # noqa: E731
nch, nsteps, n, dim_rad, lmax, L = arg(1, 512), arg(2, 10), arg(3, 50), arg(4, 41), arg(5, 25), arg(6, 2)
rng = np.random.default_rng(7)
i = np.arange(n)
v = 1.5 * np.cos(2 * np.pi * (i + 0.5) / n)
w = np.log(0.5 + 0.2 * np.cos(2 * np.pi * (i + 1.0) / n))
wrad0 = np.full(n, np.log(0.4))
lags = np.array([5.0 * (l + 1) for l in range(L)])
zeros, table = O.bessel_tables(lmax, dim_rad)
rate = O.rate_matrix(v, w, True)
counts = np.zeros((L, dim_rad, n, n))
for l, lt in enumerate(lags):
    P = np.clip(O.radial_propagator(rate, wrad0, lt, zeros, table), 0.0, None)
    counts[l] = rng.poisson(2000.0 * P)
eng = Engine(0)
ch = RadialChains(eng, v, w, True, lags, counts, table, zeros, nch)
ch.set_state(wrad0, dwrad=0.05)
t = time.time(); ll0 = ch.init_log_like(); torch.cuda.synchronize(); t_init = time.time() - t
ch.run(2, num_mc_update=100, seed=1); torch.cuda.synchronize()
t = time.time(); ch.run(nsteps, num_mc_update=100, seed=1); torch.cuda.synchronize(); el = time.time() - t
st = ch.host_state()
t = time.time(); ref = O.rad_log_like_lag(rate, wrad0, lags, counts, zeros, table); t_cpu = time.time() - t
print("radial n=%d dim_rad=%d lmax=%d L=%d: %d chains x %d steps in %.3f s -> %.1f MC steps/s (%.2f ms per step per wave of 148); "
      "init %.3f s; ll0 rel err vs oracle %.2e; oracle port %.3f s per likelihood on one core -> GPU/CPU-core %.0fx; acc %.2f status %s" % (
          n, dim_rad, lmax, L, nch, nsteps, el, nch * nsteps / el, 1e3 * el / nsteps / max(1.0, nch / 148.0), t_init,
          abs(ll0[0] - ref) / abs(ref), t_cpu, nch * nsteps / el * t_cpu, st["counters"][:, 0].mean() / (nsteps + 2),
          st["counters"][:, 3].any()))

if hasattr(eng.lib, "mcd_debug_phase_clocks"):      # -DMCD_PHASE_CLOCK build (MCDIFF_B200_LIB=...)
    import ctypes as C
    out = (C.c_longlong * 32)()
    eng.lib.mcd_debug_phase_clocks(out)             # reset
    ch.run(nsteps, num_mc_update=100, seed=1); torch.cuda.synchronize()
    eng.lib.mcd_debug_phase_clocks(out)
    names = {20: "lag products, copies to the workspace", 21: "Bessel sums + logarithms", 22: "batch diagonals, Gershgorin",
             23: "squarings of the batch", 24: "proposal, Metropolis, records", 25: "Taylor polynomials of the batch"}
    tot = float(sum(out)) or 1.0
    print("cycles per MC step (CTA 0): %.0f" % (tot / nsteps))
    for i in range(32):
        if out[i]:
            print("%2d %-42s %9.0f cycles/step %5.1f %%" % (i, names.get(i, "(chain-body phase %d)" % i), out[i] / nsteps,
                                                          100.0 * out[i] / tot))

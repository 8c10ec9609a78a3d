"""Config-4 shape: 16 independent single-lag problems (100 bins, CosinusModel 10/6) advancing concurrently on 16 CUDA
streams (mcdiff_b200.infty.run_lag_scan), then the mcdiff-infty extrapolation.
usage: prof_lagscan.py [chains per lag] [MC steps]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcdiff_b200.infty import extrapolate, run_lag_scan
from mcdiff_b200.transitions import write_Tmat_square
from oracle import mcdiff_oracle as O

nch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nmc = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
lags = [float(5 * (k + 1)) for k in range(16)]
counts, lg, edges, F, w = O.synthetic_counts(n=100, lags=lags, seed=11, total=3.0e5)
tmp = tempfile.mkdtemp()
by_lag = {}
for l, lag in enumerate(lags):
    fn = os.path.join(tmp, "transitions.nbins100.%d.pbc.dat" % int(lag))
    write_Tmat_square(counts[l], fn, lag, "pbc", edges=edges, dt=1.0, dn=int(lag))
    by_lag[lag] = fn
run_lag_scan(by_lag, os.path.join(tmp, "warm"), nmc=100, nmc_update=100.0, nchains=nch, chunk=100)   # warm-up
torch.cuda.synchronize()
t = time.time()
outs = run_lag_scan(by_lag, os.path.join(tmp, "scan"), nmc=nmc, nmc_update=1000.0, nchains=nch, chunk=1000)
torch.cuda.synchronize()
el = time.time() - t
lts, Dreal, t0 = extrapolate(outs, os.path.join(tmp, "hexd."))
print("lag scan: 16 lags x %d chains x %d MC steps in %.2f s (incl. file reading, uploads, first likelihoods, output) -> "
      "%.3e MC steps/s aggregate; extrapolated D finite: %s" % (nch, nmc, el, 16 * nch * nmc / el, bool(np.all(np.isfinite(Dreal)))))

"""Cost of one Taylor/Horner pass in the product kernel: time 148 chains x 200 MC steps with the Taylor order forced
to 1, 6, 12, 18, 24 (mcd_mc_params.taylor_m); the slope is the cycles per pass.  (Orders below the adaptive default
give wrong propagators -- timing only.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcdiff_b200 import _capi as K
from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch
from oracle import mcdiff_oracle as O

counts, lags, edges, F, w_true = O.synthetic_counts()
n = 100
vb, wb = O.cos_basis_center(n, 10), O.cos_basis_border(n, n, 6)
eng = Engine(0)
prob = DeviceProblem(eng, counts, lags, True, v_basis=vb, w_basis=wb)
nch, nsteps = 148, 200
res = []
for tm in (0, 1, 6, 12, 18, 24):
    ch = ChainBatch(prob, nch)
    ch.set_state(np.zeros(n), wb @ np.zeros(6), np.zeros(10), np.zeros(6), dv=0.0005, dw=0.0005)
    ch.init_log_like()
    mp = K.McParams(); mp.num_mc_update = 0.0; mp.min_lt = 10.0; mp.sample_every = 100; mp.refresh_every = 512; mp.seed = 1
    mp.taylor_m = tm
    ch.run(20, mp); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ch.run(nsteps, mp); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res.append((tm, ms))
    print("taylor_m %2d: %.3f ms for %d steps -> %.0f cycles per step at 1.965 GHz" % (tm, ms, nsteps, ms * 1e-3 * 1.965e9 / nsteps))
(t1, m1), (t2, m2) = res[1], res[-1]
print("slope: %.0f cycles per Horner pass" % ((m2 - m1) * 1e-3 * 1.965e9 / nsteps / (t2 - t1)))

"""--pull on the bench workload (100 bins x 4 lags, CosinusModel 10/6): MC steps/s of the general squaring engine."""
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
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for pull in (None, 0.05):
    prob = DeviceProblem(eng, counts, lags, True, v_basis=vb, w_basis=wb, pull=pull)
    ch = ChainBatch(prob, nch)
    ch.set_state(np.zeros(n), wb @ np.zeros(6), np.zeros(10), np.zeros(6), dv=0.05, dw=0.05)
    ch.init_log_like()
    mp = K.McParams(); mp.num_mc_update = 1000.0; mp.min_lt = 10.0; mp.sample_every = 100; mp.refresh_every = 512; mp.seed = 1
    ch.run(1000, mp); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ch.run(200, mp); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = ch.host_state()
    print("pull %s: %d chains x 200 steps in %.1f ms -> %.3e MC steps/s (squarings/step %.2f, status %s)" % (
        pull, nch, ms, nch * 200 / (ms * 1e-3), st["counters"][:, 7].sum() / st["counters"][:, 5].sum(), bool(st["counters"][:, 6].any())))

"""MC steps/s of the 1-D kernel versus the number of bins (4 lags 10..40, CosinusModel 10/6, 1184 chains = 8 waves)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcdiff_b200 import _capi as K
from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch
from oracle import mcdiff_oracle as O

eng = Engine(0)
nch = 1184
for n in (24, 32, 48, 56, 64, 80, 100, 104, 120, 160, 200, 250, 312, 400, 508):
    counts, lags, edges, F, w_true = O.synthetic_counts(n=n, lags=(10.0, 20.0, 30.0, 40.0), total=3.0e5, dz=50.0 / n)
    vb, wb = O.cos_basis_center(n, 10), O.cos_basis_border(n, n, 6)
    prob = DeviceProblem(eng, counts, lags, True, v_basis=vb, w_basis=wb)
    nch = 296 if n > 208 else 1184
    ch = ChainBatch(prob, nch)
    wc0 = np.zeros(6); wc0[0] = float(np.mean(w_true))
    ch.set_state(np.zeros(n), wb @ wc0, np.zeros(10), wc0, dv=0.05, dw=0.05)
    ch.init_log_like()
    mp = K.McParams(); mp.num_mc_update = 1000.0; mp.min_lt = 10.0; mp.sample_every = 100; mp.refresh_every = 512; mp.seed = 1
    nst = 20 if n > 208 else (100 if n > 104 else 200)
    ch.run(nst, mp); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ch.run(nst, mp); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = ch.host_state()
    fl = 9.0 * n ** 3 + 4 * (2.0 * n ** 3 + 3.0 * n ** 2)
    print("n = %3d: %.3e MC steps/s, %.1f algorithmic TFLOP/s, squarings/step %.2f" % (
        n, nch * nst / (ms * 1e-3), fl * nch * nst / (ms * 1e-3) / 1e12, st["counters"][:, 7].sum() / st["counters"][:, 5].sum()))

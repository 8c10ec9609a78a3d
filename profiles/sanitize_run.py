"""Tiny workload touching every kernel path, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcdiff_b200 import _capi as K
from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch, RadialChains, rad_loglike
from oracle import mcdiff_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "loglike_golden.npz"))
Rd = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "radial_golden.npz"))
eng = Engine(0)
n = 100
vb, wb = O.cos_basis_center(n, 6), O.cos_basis_border(n, n, 4)
for lags, t0 in ((G["P_lags"], 0), ([10.0, 20.0, 35.0, 40.0], 1)):
    prob = DeviceProblem(eng, G["P_counts"], lags, True, v_basis=vb, w_basis=wb)
    for method in (1, 2, 3):
        ll, _, st = prob.loglike(G["rand_v"][:2], G["rand_wP"][:2], method=method, want_P=(method == 2))
        ch = ChainBatch(prob, 3)
        wc0 = np.zeros(4); wc0[0] = G["P_flat_w"][0]
        ch.set_state(np.zeros(n), wb @ wc0, np.zeros(6), wc0, dv=0.3, dw=0.3, dtimezero=0.5)
        ch.init_log_like()
        mp = K.McParams()
        mp.num_mc_update = 2.0; mp.min_lt = 10.0; mp.sample_every = 2; mp.refresh_every = 2; mp.seed = 5
        mp.method = method; mp.move_timezero = t0
        ch.run(4, mp, record=True, nsamples_cap=2)
        ch.run(3, mp)
        torch.cuda.synchronize()
        print("method", method, "t0", t0, ll, ch.host_state()["scal"][:, 0])
pS = DeviceProblem(eng, G["S_counts"], G["S_lags"], True)            # n = 37 padding path
print(pS.loglike(G["S_v"][:2], G["S_w"][:2], method=2)[0], pS.loglike(G["S_v"][:2], G["S_w"][:2], method=1)[0])
pC = DeviceProblem(eng, G["C_counts"], G["C_lags"], False)           # nopbc
print(pC.loglike(G["C_v"][:2], G["C_w"][:2], method=2)[0])
ll, P, st = rad_loglike(eng, Rd["v"], Rd["w"], True, Rd["lags"], Rd["counts"], Rd["table"], Rd["zeros"], Rd["wrad"][:2])
rc = RadialChains(eng, Rd["v"], Rd["w"], True, Rd["lags"], Rd["counts"], Rd["table"], Rd["zeros"], 2)
rc.set_state(Rd["wrad"][0], dwrad=0.3)
rc.init_log_like()
rc.run(3, num_mc_update=2.0, seed=3, record=True, nsamples_cap=1)
torch.cuda.synchronize()
print("radial", ll, rc.host_state()["scal"][:, 0])
print("ok")

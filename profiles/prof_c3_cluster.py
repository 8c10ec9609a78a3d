"""Config 3 (200 bins x 8 lags, blocked engine): one SM per chain against a two-CTA cluster per chain
(MCDIFF_B200_CLUSTER=1), 74 .. 512 chains x 20 MC steps, CUDA events."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mcdiff_b200 import _capi as K
from mcdiff_b200.engine import ChainBatch, DeviceProblem, Engine
from mcdiff_b200.model import CosinusModel
from mcdiff_b200.transitions import Transitions
eng = Engine(0)
tmp = tempfile.mkdtemp()
lags = (5.0, 10.0, 15.0, 20.0, 25.0, 30.0, 35.0, 40.0)
c3, e3 = bench.synthetic_counts(200, lags, seed=3)
bench.write_square_files(tmp, c3, lags, e3)
data = Transitions([os.path.join(tmp, "transitions.nbins200.%d.pbc.dat" % int(l)) for l in lags])
model = CosinusModel(data, 1.0, 12, 8)
prob = DeviceProblem(eng, np.asarray(data.list_trans).astype(np.int64), data.list_lt, True, v_basis=model.v_basis, w_basis=model.w_basis)
mp = K.McParams(); mp.num_mc_update = 1000.0; mp.min_lt = 5.0; mp.sample_every = 100; mp.refresh_every = 512; mp.seed = 1; mp.method = 0
for cl in (0, 1):
    os.environ["MCDIFF_B200_CLUSTER"] = "1" if cl else "0"
    for nch in (74, 148, 296, 512):
        ch = ChainBatch(prob, nch)
        ch.set_state(model.v, model.w, model.v_coeff, model.w_coeff, dv=0.05, dw=0.05)
        ch.init_log_like()
        ch.run(20, mp); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ch.run(20, mp); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("cluster=%d chains=%4d %.2f ms / 20 steps -> %.3e MC steps/s" % (cl, nch, ms, nch * 20 / (ms * 1e-3)))

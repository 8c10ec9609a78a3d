"""Headline workload (100 bins x 4 lags, 1036 = 7 x 148 chains x 200 MC steps per launch) timed with CUDA events for the
library named by MCDIFF_B200_LIB (code-size variants: profiles/_variants/*.so, built by the recipe in DESIGN.md):
    MCDIFF_B200_LIB=profiles/_variants/lib_v1.so python profiles/prof_headline.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mcdiff_b200 import _capi as K
from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch

import tempfile
import numpy as np
from mcdiff_b200.model import CosinusModel
from mcdiff_b200.transitions import Transitions
eng = Engine(0)
tmp = tempfile.mkdtemp(prefix="mcd_prof_")
bench.headline_files(tmp)
data = Transitions([os.path.join(tmp, "transitions.nbins%d.%d.pbc.dat" % (bench.N_BINS, int(l))) for l in bench.LAGS])
model = CosinusModel(data, 1.0, bench.NCOS_F, bench.NCOS_D)
prob = DeviceProblem(eng, np.asarray(data.list_trans).astype(np.int64), data.list_lt, True,
                     v_basis=model.v_basis, w_basis=model.w_basis)
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 7 * 148
ch = ChainBatch(prob, nch)
ch.set_state(model.v, model.w, model.v_coeff, model.w_coeff, dv=bench.DV, dw=bench.DW)
ch.init_log_like()
mp = K.McParams()
mp.num_mc_update = bench.NMC_UPDATE; mp.min_lt = 10.0; mp.sample_every = 100; mp.refresh_every = 512; mp.seed = 1
for _ in range(3):
    ch.run(bench.MC_PER_LAUNCH, mp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ch.run(bench.MC_PER_LAUNCH, mp)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
st = ch.host_state()
print("%s: %d chains x %d steps: %.2f ms/launch -> %.4e MC steps/s, ll[0] = %.10g" % (
    os.path.basename(K.LIB_PATH), nch, bench.MC_PER_LAUNCH, ms, nch * bench.MC_PER_LAUNCH / (ms * 1e-3), st["scal"][0, 0]))

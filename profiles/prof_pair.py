"""One chain per SM versus two chains per SM: headline workload, 148 and 296 chains x 100 MC steps, full-storage kernel
and packed kernel (MCDIFF_B200_PACKED=1), timed with CUDA events.  Says what a chain costs alone and what the overlap of
two resident chains buys."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mcdiff_b200 import _capi as K
from mcdiff_b200.engine import ChainBatch, DeviceProblem, Engine
from mcdiff_b200.model import CosinusModel
from mcdiff_b200.transitions import Transitions

eng = Engine(0)
tmp = tempfile.mkdtemp(prefix="mcd_pair_")
bench.headline_files(tmp)
files = [os.path.join(tmp, "transitions.nbins%d.%d.pbc.dat" % (bench.N_BINS, int(l))) for l in bench.LAGS]
data = Transitions(files)
model = CosinusModel(data, 1.0, bench.NCOS_F, bench.NCOS_D)
prob = DeviceProblem(eng, np.asarray(data.list_trans).astype(np.int64), data.list_lt, True,
                     v_basis=model.v_basis, w_basis=model.w_basis)
mp = K.McParams()
mp.num_mc_update = 1000.0; mp.min_lt = 10.0; mp.sample_every = 100; mp.refresh_every = 512; mp.seed = 1; mp.method = 0
for packed in (0, 1):
    os.environ["MCDIFF_B200_PACKED"] = "1" if packed else "0"
    for nch in (148, 296, 592, 1024):
        ch = ChainBatch(prob, nch)
        ch.set_state(model.v, model.w, model.v_coeff, model.w_coeff, dv=bench.DV, dw=bench.DW)
        ch.init_log_like()
        ch.run(300, mp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ch.run(100, mp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("packed=%d chains=%4d  %.2f ms / 100 steps  -> %.3e MC steps/s" % (packed, nch, ms, nch * 100 / (ms * 1e-3)))

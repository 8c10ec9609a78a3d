"""One launch of the bench configuration (1024 chains x 200 MC steps) after a short warm-up,
for `ncu --set full -k regex:mcd_mc_run -s 1 -c 1` (DRAM traffic per launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mcdiff_b200 import _capi as K
from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch

eng = Engine(0)
wl = bench.workload()
prob = DeviceProblem(eng, wl["counts"], wl["lags"], True, v_basis=wl["vb"], w_basis=wl["wb"])
ch = ChainBatch(prob, bench.CHAINS_PER_GPU)
ch.set_state(wl["v0"], wl["w0"], wl["vc0"], wl["wc0"], dv=bench.DV, dw=bench.DW)
ch.init_log_like()
mp = K.McParams()
mp.num_mc_update = bench.NMC_UPDATE; mp.min_lt = 10.0; mp.sample_every = 100; mp.refresh_every = 512; mp.seed = 1
ch.run(20, mp)                                   # launch 0: warm-up (skipped by -s 1)
ch.run(bench.MC_PER_LAUNCH, mp, nsamples_cap=2)  # launch 1: the bench launch
torch.cuda.synchronize()
print("ok")

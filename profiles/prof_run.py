"""Small fixed workload for ncu: 296 chains x 20 MC steps of the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mcdiff_b200 import _capi as K
from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch

method = {"eigen": 1, "squaring": 2, "squaring_dfma": 3}[sys.argv[1] if len(sys.argv) > 1 else "squaring"]
eng = Engine(0)
wl = bench.workload()
prob = DeviceProblem(eng, wl["counts"], wl["lags"], True, v_basis=wl["vb"], w_basis=wl["wb"])
ch = ChainBatch(prob, 296)
ch.set_state(wl["v0"], wl["w0"], wl["vc0"], wl["wc0"], dv=0.05, dw=0.05)
ch.init_log_like()
mp = K.McParams()
mp.num_mc_update = 1000.0; mp.min_lt = 10.0; mp.sample_every = 100; mp.refresh_every = 512; mp.seed = 1; mp.method = method
for _ in range(3):
    ch.run(20, mp)
torch.cuda.synchronize()
print("ok", ch.host_state()["scal"][:2, 0])

"""Small fixed workloads for ncu (one wave of CTAs, a few launches; capture the last one):
  python profiles/prof_run.py main   [method]   296 chains x 20 MC steps of the bench headline (100 bins x 4 lags)
  python profiles/prof_run.py c3                148 chains x 10 MC steps of config 3 (200 bins x 8 lags, blocked engine)
  python profiles/prof_run.py radial [chains]   592 chains (4 per SM) x 20 MC steps of config 5 (n=30, dim_rad=41, lmax=20)
The workloads are the ones bench.py times (same generators, same models)."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mcdiff_b200 import _capi as K
from mcdiff_b200 import engine as E
from mcdiff_b200.engine import ChainBatch, DeviceProblem, Engine, RadialChains, RadialProblem
from mcdiff_b200.model import CosinusModel, RadCosinusModel
from mcdiff_b200.transitions import RadTransitions, Transitions, write_Tmat_cube

what = sys.argv[1] if len(sys.argv) > 1 else "main"
eng = Engine(0)
tmp = tempfile.mkdtemp(prefix="mcd_prof_")


def params(lags, method=0):
    mp = K.McParams()
    mp.num_mc_update = 1000.0
    mp.min_lt = float(min(lags))
    mp.sample_every = 100
    mp.refresh_every = 512
    mp.seed = 1
    mp.method = method
    return mp


if what == "main":
    method = {"auto": 0, "eigen": 1, "squaring": 2, "squaring_dfma": 3}[sys.argv[2] if len(sys.argv) > 2 else "auto"]
    bench.headline_files(tmp)
    files = [os.path.join(tmp, "transitions.nbins%d.%d.pbc.dat" % (bench.N_BINS, int(l))) for l in bench.LAGS]
    data = Transitions(files)
    model = CosinusModel(data, 1.0, bench.NCOS_F, bench.NCOS_D)
    prob = DeviceProblem(eng, np.asarray(data.list_trans).astype(np.int64), data.list_lt, True,
                         v_basis=model.v_basis, w_basis=model.w_basis)
    ch = ChainBatch(prob, 296)
    ch.set_state(model.v, model.w, model.v_coeff, model.w_coeff, dv=bench.DV, dw=bench.DW)
    ch.init_log_like()
    mp = params(bench.LAGS, method)
    ch.run(400, mp)                          # towards the steady state the bench measures
    for _ in range(3):
        ch.run(20, mp)
    torch.cuda.synchronize()
    print("ok", ch.host_state()["scal"][:2, 0])
elif what == "c3":
    lags = (5.0, 10.0, 15.0, 20.0, 25.0, 30.0, 35.0, 40.0)
    c3, e3 = bench.synthetic_counts(200, lags, seed=3)
    bench.write_square_files(tmp, c3, lags, e3)
    data = Transitions([os.path.join(tmp, "transitions.nbins200.%d.pbc.dat" % int(l)) for l in lags])
    model = CosinusModel(data, 1.0, 12, 8)
    prob = DeviceProblem(eng, np.asarray(data.list_trans).astype(np.int64), data.list_lt, True,
                         v_basis=model.v_basis, w_basis=model.w_basis)
    ch = ChainBatch(prob, 148)
    ch.set_state(model.v, model.w, model.v_coeff, model.w_coeff, dv=bench.DV, dw=bench.DW)
    ch.init_log_like()
    mp = params(lags)
    for _ in range(3):
        ch.run(10, mp)
    torch.cuda.synchronize()
    print("ok", ch.host_state()["scal"][:2, 0])
else:
    from scipy import special
    RN, RDIM, RLMAX, RLAGS = 30, 41, 20, (5.0, 10.0)
    v = np.zeros(RN)
    w = np.full(RN, -np.log(0.25))
    wrad = np.log((0.4 + 0.15 * np.cos(2 * np.pi * (np.arange(RN) + 0.5) / RN)) / 0.25)
    zeros = special.jn_zeros(0, RLMAX)
    r = np.arange(RDIM, dtype=np.float64) + 0.5
    bess = np.array([2 * r * special.j0(r / RDIM * z) / special.j1(z) ** 2 / RDIM ** 2 for z in zeros])
    p0 = RadialProblem(eng, v, w, True, RLAGS, np.zeros((len(RLAGS), RDIM, RN, RN)), bess, zeros)
    _, P, _ = p0.loglike_device(eng.to_device(wrad[None, :], torch.float64), want_P=True)
    P = np.clip(E.to_host(P)[0], 0.0, None)
    p0.close()
    cube = np.random.default_rng(0).poisson(2000.0 * P * np.exp(-v)[None, None, None, :])
    fns = []
    for l, lag in enumerate(RLAGS):
        fn = os.path.join(tmp, "rad.%d.dat" % int(lag))
        write_Tmat_cube(cube[l].astype(int), fn, float(lag), "pbc", edges=np.linspace(-7.5, 7.5, RN + 1),
                        redges=np.arange(RDIM) * 0.5, dt=1.0, dn=int(lag))
        fns.append(fn)
    rdata = RadTransitions(fns)
    rmodel = RadCosinusModel(rdata, 1.0, 0, 0, 0, 4)
    rr = np.arange(rmodel.dim_rad, dtype=np.float64) + 0.5
    bess = np.array([2 * rr * special.j0(rr / rmodel.dim_rad * z) / special.j1(z) ** 2 / rmodel.dim_rad ** 2 for z in zeros])
    rch = RadialChains(eng, rmodel.v, rmodel.w, True, rdata.list_lt, rdata.list_trans, bess, zeros,
                       int(sys.argv[2]) if len(sys.argv) > 2 else 592,      # 592 = 148 SMs x 4 resident chains
                       wrad_basis=rmodel.wrad_basis)
    rch.set_state(rmodel.wrad, rmodel.wrad_coeff, dwrad=0.05)
    rch.init_log_like()
    for _ in range(3):
        rch.run(20, num_mc_update=1000.0, seed=1)
    torch.cuda.synchronize()
    print("ok", rch.host_state()["scal"][:2, 0])

"""Config-3 shape (n = 200, 8 lags, ncosF = 12, ncosD = 8): time per MC step.
usage: prof_n200.py [nchains] [nsteps] [method: 1 eigen on global matrices, 2 blocked tensor-core squaring]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcdiff_b200 import _capi as K
from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch
from oracle import mcdiff_oracle as O

n = 200
lags = (5.0, 10.0, 15.0, 20.0, 25.0, 30.0, 35.0, 40.0)
counts, lags, edges, F, w_true = O.synthetic_counts(n=n, lags=lags, seed=3, total=1.0e6)
vb, wb = O.cos_basis_center(n, 12), O.cos_basis_border(n, n, 8)
eng = Engine(0)
prob = DeviceProblem(eng, counts, lags, True, v_basis=vb, w_basis=wb)
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 148
ch = ChainBatch(prob, nch)
wc0 = np.zeros(8); wc0[0] = -np.log(0.25)
ch.set_state(np.zeros(n), wb @ wc0, np.zeros(12), wc0, dv=0.05, dw=0.05)
t = time.time(); ch.init_log_like(); torch.cuda.synchronize(); print("init_log_like %.3f s" % (time.time() - t))
mp = K.McParams(); mp.num_mc_update = 1000.0; mp.min_lt = 5.0; mp.sample_every = 100; mp.refresh_every = 512; mp.seed = 1
mp.method = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ch.run(2, mp); torch.cuda.synchronize()
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
t = time.time(); ch.run(nsteps, mp); torch.cuda.synchronize(); el = time.time() - t
st = ch.host_state()
print("method %d" % mp.method, "n=200 L=8: %d chains x %d steps in %.3f s -> %.1f MC steps/s, %.2f ms per step per wave, sweeps/step %.2f, status %s, in_smem %s squaring_ok %s" % (
    nch, nsteps, el, nch * nsteps / el, 1e3 * el / nsteps, st["counters"][:, 7].sum() / st["counters"][:, 5].sum(), st["counters"][:, 6].any(), prob.in_smem, prob.squaring_ok))

if hasattr(eng.lib, "mcd_debug_phase_clocks"):      # -DMCD_PHASE_CLOCK build (MCDIFF_B200_LIB=...)
    import ctypes as C
    out = (C.c_longlong * 32)()
    eng.lib.mcd_debug_phase_clocks(out)
    ch.run(nsteps, mp); torch.cuda.synchronize()
    eng.lib.mcd_debug_phase_clocks(out)
    names = {0: "proposal", 1: "trial profile", 2: "build_S + reductions", 4: "Taylor/Horner (band, shared memory) + expansion",
             25: "panel staging (cp.async)", 5: "block prologue (task, partial sums, counts)", 6: "m8n8k4 k loop",
             7: "likelihood epilogue", 8: "result stores (+ transpose)", 9: "barrier wait after block", 11: "final reduction",
             12: "Metropolis", 13: "commit, records"}
    tot = float(sum(out))
    print("cycles per MC step (CTA 0): %.0f" % (tot / nsteps))
    for i in range(32):
        if out[i]:
            print("%2d %-50s %9.0f cycles/step %5.1f %%" % (i, names.get(i, "(phase %d)" % i), out[i] / nsteps, 100.0 * out[i] / tot))

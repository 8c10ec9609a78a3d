"""Per-phase cycle counts of the MC kernel (bench workload) from a -DMCD_PHASE_CLOCK build.
Build:  nvcc <flags of __graft_entry__> -DMCD_PHASE_CLOCK mcdiff_b200/csrc/mcd_lib.cu mcdiff_b200/csrc/mcd_small.cu -o mcdiff_b200/libmcdiff_b200_clk.so
Run:    MCDIFF_B200_LIB=mcdiff_b200/libmcdiff_b200_clk.so python profiles/prof_phases.py
Thread 0 of CTA 0 stamps clock64() at phase boundaries; the instrumented library is never the product."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcdiff_b200 import _capi as K
from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch
from oracle import mcdiff_oracle as O

NAMES = {0: "proposal (Philox, thread 0)", 1: "trial profile (basis . coeff)", 2: "build_S, Gershgorin, reductions",
         3: "zero work matrices", 4: "Taylor/Horner passes", 5: "product prologue (task table, count prefetch)",
         6: "m8n8k4 k loop", 7: "likelihood epilogue", 8: "result stores (+ mirror)", 9: "barrier wait after product",
         10: "staged store phase (in-place products)", 11: "final reduction", 12: "Metropolis (thread 0)",
         13: "commit, records, adaptation", 14: "Taylor sweeps (thread 0)", 15: "Taylor band stores (thread 0)",
         16: "squaring count, Taylor order, h", 17: "between products (loop control, call)", 18: "task decode", 19: "Taylor sweep prologue (loads, centre exchange)", 20: "Taylor sweep chunks"}
counts, lags, edges, F, w_true = O.synthetic_counts()
n = 100
vb, wb = O.cos_basis_center(n, 10), O.cos_basis_border(n, n, 6)
eng = Engine(0)
prob = DeviceProblem(eng, counts, lags, True, v_basis=vb, w_basis=wb)
nch, nsteps = 148, 200
ch = ChainBatch(prob, nch)
wc0 = np.zeros(6)
ch.set_state(np.zeros(n), wb @ wc0, np.zeros(10), wc0, dv=0.05, dw=0.05)
ch.init_log_like()
mp = K.McParams(); mp.num_mc_update = 1000.0; mp.min_lt = 10.0; mp.sample_every = 100; mp.refresh_every = 512; mp.seed = 1
ch.run(3000, mp); torch.cuda.synchronize()          # reach the steady state the bench measures (about 7.6 squarings, m = 12)
out = (C.c_longlong * 32)()
eng.lib.mcd_debug_phase_clocks(out)          # reset after warm-up
ch.run(nsteps, mp); torch.cuda.synchronize()
eng.lib.mcd_debug_phase_clocks(out)
tot = float(sum(out))
print("cycles per MC step (CTA 0, thread 0): %.0f" % (tot / nsteps))
for i in range(21):
    print("%2d %-48s %9.0f cycles/step  %5.1f %%" % (i, NAMES[i], out[i] / nsteps, 100.0 * out[i] / tot))

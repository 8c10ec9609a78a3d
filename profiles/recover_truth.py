"""End-to-end sanity at scale: counts are drawn from a known F(z), D(z) (oracle.synthetic_counts, 1e6 transitions per
lag); 1024 chains x 20 000 MC steps of the CosinusModel 10/6 sampler must recover the truth.
Prints the RMS deviation of the posterior-mean F (up to a constant) and ln D from the truth and the
between-chain spread."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcdiff_b200 import _capi as K
from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch
from oracle import mcdiff_oracle as O

counts, lags, edges, F_true, w_true = O.synthetic_counts()
n = 100
ncF, ncD = 10, 6
vb, wb = O.cos_basis_center(n, ncF), O.cos_basis_border(n, n, ncD)
eng = Engine(0)
prob = DeviceProblem(eng, counts, lags, True, v_basis=vb, w_basis=wb)
nch, nsteps = 1024, 20000
ch = ChainBatch(prob, nch)
wc0 = np.zeros(ncD); wc0[0] = float(np.mean(w_true))
ch.set_state(np.zeros(n), wb @ wc0, np.zeros(ncF), wc0, dv=0.05, dw=0.05)
ll0 = ch.init_log_like()
mp = K.McParams(); mp.num_mc_update = 1000.0; mp.min_lt = 10.0; mp.sample_every = 100; mp.refresh_every = 512; mp.seed = 2024
t = time.time()
rows = []
for chunk in range(nsteps // 1000):
    _, _, smp = ch.run(1000, mp, nsamples_cap=10)
    rows.append(smp.cpu().numpy())
torch.cuda.synchronize()
el = time.time() - t
S = np.concatenate(rows, axis=1)                       # [chain][sample][6 + ncF + ncD]
half = S[:, S.shape[1] // 2:, :]
vc = half[:, :, 6:6 + ncF]; wc = half[:, :, 6 + ncF:]
Fm = (vc.mean(axis=(0, 1)) @ vb.T); Wm = (wc.mean(axis=(0, 1)) @ wb.T)
# the truth is not exactly representable by 10/6 cosines: compare with its projection too
def proj(basis, y):
    c, *_ = np.linalg.lstsq(basis, y, rcond=None); return basis @ c
dF = (Fm - Fm.mean()) - (F_true - F_true.mean())
dW = Wm - w_true
dFp = (Fm - Fm.mean()) - (proj(vb, F_true) - proj(vb, F_true).mean())
dWp = Wm - proj(wb, w_true)
chainF = (vc.mean(axis=1) @ vb.T); chainF -= chainF.mean(axis=1, keepdims=True)
st = ch.host_state()
print("1024 chains x %d steps in %.1f s (%.2e MC steps/s incl. sample read-back); log L: start %.1f -> mean %.1f"
      % (nsteps, el, nch * nsteps / el, ll0[0], st["scal"][:, 0].mean()))
print("posterior mean vs truth: RMS dF = %.4f kT (%.4f vs the truth's projection on the basis), RMS d lnD = %.4f (%.4f); "
      "between-chain std of F: %.4f kT; acceptance %.2f; chains with status: %d"
      % (np.sqrt(np.mean(dF ** 2)), np.sqrt(np.mean(dFp ** 2)), np.sqrt(np.mean(dW ** 2)), np.sqrt(np.mean(dWp ** 2)),
         chainF.std(axis=0).mean(), (st["counters"][:, 0] + st["counters"][:, 1]).mean() / nsteps, int((st["counters"][:, 6] != 0).sum())))

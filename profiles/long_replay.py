"""Long accept/reject parity: 8 chains x 2000 MC steps on device Philox streams (bench workload, time-offset moves on)
against the oracle replay of the same tapes on the host cores."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from concurrent.futures import ProcessPoolExecutor
from mcdiff_b200 import _capi as K, philox
from oracle import mcdiff_oracle as O

NCH, NSTEPS, SEED, N = 8, 2000, 4242, 100
counts, lags, edges, F, w_true = O.synthetic_counts()
vb, wb = O.cos_basis_center(N, 10), O.cos_basis_border(N, N, 6)
wc0 = np.zeros(6); wc0[0] = float(np.mean(w_true))


def replay(c):
    cfg = O.ChainConfig(counts, lags, True, v_basis=vb, w_basis=wb, dv=0.05, dw=0.05, dtimezero=0.2, temp=1.0, nmc=NSTEPS,
                        num_mc_update=500.0, move_timezero=True)
    tu, ti = philox.chain_tape(SEED, c, 0, NSTEPS, N, N, 10, 6)
    out = O.replay(cfg, np.zeros(N), wb @ wc0, tu, ti, v_coeff0=np.zeros(10), w_coeff0=wc0)
    return out["accepted"], out["accepted_t0"], out["ll"], np.nanmin(np.abs(out["margin"])) if "margin" in out else np.nan


if __name__ == "__main__":
    from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch
    eng = Engine(0)
    prob = DeviceProblem(eng, counts, lags, True, v_basis=vb, w_basis=wb)
    ch = ChainBatch(prob, NCH)
    ch.set_state(np.zeros(N), wb @ wc0, np.zeros(10), wc0, dv=0.05, dw=0.05, dtimezero=0.2)
    ch.init_log_like()
    mp = K.McParams(); mp.num_mc_update = 500.0; mp.min_lt = 10.0; mp.move_timezero = 1; mp.sample_every = 100
    mp.refresh_every = 512; mp.seed = SEED
    decs = []
    for chunk in range(4):
        d, _, _ = ch.run(NSTEPS // 4, mp, record=True)
        decs.append(d.cpu().numpy())
    dec = np.concatenate(decs, axis=1)
    st = ch.host_state()
    t = time.time()
    with ProcessPoolExecutor(max_workers=min(NCH, os.cpu_count() or 1)) as ex:
        refs = list(ex.map(replay, range(NCH)))
    same = sum(int(np.array_equal(dec[c] & 1, refs[c][0]) and np.array_equal((dec[c] >> 1) & 1, refs[c][1])) for c in range(NCH))
    relll = max(abs(st["scal"][c][0] - refs[c][2]) / abs(refs[c][2]) for c in range(NCH))
    print("%d chains x %d steps (2 decisions per step): %d/%d chains identical to the oracle replay, final log L rel err %.1e, "
          "accepted %d + %d (t0); oracle replay took %.0f s on the host" % (
              NCH, NSTEPS, same, NCH, relll, int((dec & 1).sum()), int(((dec >> 1) & 1).sum()), time.time() - t))

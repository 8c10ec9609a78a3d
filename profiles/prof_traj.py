"""Trajectory -> counts (mcd_traj_counts) on the bench workload: 2^28 coordinates, 100 bins, 4 shifts.
  python profiles/prof_traj.py            CUDA-event time per call for 4 / 1 / 2 / 11 shifts and one long trajectory
  python profiles/prof_traj.py once       three calls of the default configuration (for ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcdiff_b200 import extract as X
from mcdiff_b200.engine import Engine

eng = Engine(0)
T, M, nbins, zpbc = 1 << 19, 512, 100, 50.0
shifts = [10, 20, 30, 40]
g = torch.Generator(device="cuda").manual_seed(11)
z = torch.cumsum(torch.randn((T, M), generator=g, device="cuda", dtype=torch.float64) * 0.35, dim=0)
edges = eng.to_device(np.arange(-nbins * (zpbc / nbins) / 2, (nbins + 1) * (zpbc / nbins) / 2., zpbc / nbins), torch.float64)
out = torch.zeros((len(shifts), nbins + 2, nbins + 2), dtype=torch.int64, device="cuda")


def timed(nrep=5, sh=shifts, zz=z):
    o = out if len(sh) == len(shifts) else torch.zeros((len(sh), nbins + 2, nbins + 2), dtype=torch.int64, device="cuda")
    for _ in range(2):
        X.transition_counts_device(zz, edges, sh, period=zpbc, engine=eng, out=o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nrep):
        X.transition_counts_device(zz, edges, sh, period=zpbc, engine=eng, out=o)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / nrep


if len(sys.argv) > 1 and sys.argv[1] == "once":
    for _ in range(3):
        X.transition_counts_device(z, edges, shifts, period=zpbc, engine=eng, out=out)
    torch.cuda.synchronize()
    print("ok", int(out.sum()))
    sys.exit(0)

nbytes = 8.0 * T * M
ms = timed()
print("shifts= 4  %.3f ms/call  %.0f GB/s algorithmic" % (ms, nbytes / ms / 1e6))
for sh in ([10], [10, 20], [1, 2, 3, 4, 5, 10, 15, 20, 30, 40, 50]):
    ms = timed(sh=sh)
    print("shifts=%2d  %.3f ms/call  %.0f GB/s algorithmic" % (len(sh), ms, nbytes / ms / 1e6))
z1 = z.t().contiguous().reshape(-1)[: 1 << 26]           # ONE long trajectory: lanes of a warp are consecutive frames
ms = timed(zz=z1)
print("single trajectory of 2^26 frames: %.3f ms/call  %.0f GB/s algorithmic" % (ms, 8.0 * (1 << 26) / ms / 1e6))

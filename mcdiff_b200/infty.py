"""Lag-time scan (BASELINE config 4) and extrapolation to infinite lag time.

The reference workflow (examples/multipleruns.sh + scripts/mcdiff-infty) runs one single-lag
`run-mcdiff` per lag time, one after the other, then fits `D_i(lt) = A_i + B_i/lt` per bin
(scripts/mcdiff-infty:82-165) and writes `Dreal.dat`, `t0.dat`.  Here the independent problems run
CONCURRENTLY on one GPU, each on its own CUDA stream (one CTA per chain, so 16 problems x 256 chains
fill the 148 SMs), and the fit is the closed form of the same least-squares problem.
"""
from __future__ import annotations

import os

import numpy as np

from .outreading import read_many_profiles


def extract_lagtimes_from_filenames(list_filenames, dtc=1.0):
    """Lag time = third dot-separated field of the file name, e.g. out.nbins100.20.pbc.dat
    (scripts/mcdiff-infty:33-46)."""
    return np.array([int(os.path.basename(fn).split(".")[2]) * dtc for fn in list_filenames])


def fit_line_inverse(lts, D):
    """Least-squares fit D[i, l] = A[i] + B[i] / lts[l] for every bin i.

    The reference builds one (nbins*nlt) x (2*nbins) design matrix and calls np.linalg.lstsq
    (scripts/mcdiff-infty:99-117); that system is block diagonal, one 2-parameter line fit per bin,
    solved here in closed form.  Returns (Dreal = A, t0 = B / A)."""
    lts = np.asarray(lts, dtype=np.float64)
    D = np.asarray(D, dtype=np.float64)
    x = 1.0 / lts
    n = len(lts)
    sx, sxx = x.sum(), (x * x).sum()
    sy, sxy = D.sum(axis=1), (D * x[None, :]).sum(axis=1)
    det = n * sxx - sx * sx
    A = (sxx * sy - sx * sxy) / det
    B = (n * sxy - sx * sy) / det
    return A, B / A


def print_profile(filename, v, edges, startline=None):
    """scripts/mcdiff-infty:52-59."""
    with open(filename, "w") as f:
        if startline is not None:
            print(startline, file=f)
        for i in range(len(v)):
            print("%8d %16.8e %16.8e  %16.8f" % (i, edges[i], edges[i + 1], v[i]), file=f)
        print("=" * 10, file=f)


def print_profile_2vecs(filename, v, d, edges, startline=None):
    """scripts/mcdiff-infty:62-70."""
    with open(filename, "w") as f:
        if startline is not None:
            print(startline, file=f)
        for i in range(len(v)):
            print("%8d %16.8e %16.8e  %16.8f  %16.8f" % (i, edges[i], edges[i + 1], v[i], d[i]), file=f)
        print("=" * 10, file=f)


def extrapolate(list_filenames, prefix, dtc=1.0):
    """`mcdiff-infty` without the figures: reads the final `.dat` of every lag, fits, writes
    <prefix>Dreal.dat and <prefix>t0.dat (scripts/mcdiff-infty:228-276)."""
    lts = extract_lagtimes_from_filenames(list_filenames, dtc)
    F, D, E, _, _ = read_many_profiles(list_filenames)
    F, D, E = np.array(F).T, np.array(D).T, np.array(E).T
    Dreal, t0 = fit_line_inverse(lts, D)
    print_profile_2vecs(prefix + "Dreal.dat", F[:, 0], Dreal, E[:, 0], startline="   index  bin-str  bin-end")
    print_profile(prefix + "t0.dat", t0, E[:, 0])
    return lts, Dreal, t0


def run_lag_scan(files_by_lag, outdir, *, pbc=True, model="CosinusModel", ncosF=10, ncosD=6, dv=0.05, dw=0.05,
                 D0=1.0, temp=1.0, temp_end=1.0, nmc=1000, nmc_update=100.0, seed=1, nchains=256, device=None,
                 chunk=1000, nbins_tag=None):
    """One independent single-lag problem per entry of `files_by_lag` ({lag: filename}), all advancing
    concurrently on separate CUDA streams; writes out.nbins<N>.<lag>.pbc.dat (chain 0) per lag, in the
    naming mcdiff-infty parses.  Returns the list of output files (ascending lag)."""
    import torch
    from .MCState import MCState
    from .log import Logger
    from .mc import do_mc_cycles  # noqa: F401  (same building blocks, driven per stream below)
    from .transitions import Transitions

    os.makedirs(outdir, exist_ok=True)
    lags = sorted(files_by_lag)
    runs = []
    for k, lag in enumerate(lags):
        MC = MCState(pbc, -1, nchains=nchains, device=device, seed=seed + k)
        MC.set_MC_params(dv, dw, 0.5, D0, 0.1, temp, nmc, nmc_update, False, -1, temp_end=temp_end)
        data = Transitions([files_by_lag[lag]])
        MC.set_model(model, data, ncosF, ncosD, 0, None)
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):                # uploads and launches of this problem share one stream
            MC._ensure_device()
        runs.append((lag, MC, stream))
    for lag, MC, stream in runs:                       # first likelihoods, one stream each
        with torch.cuda.stream(stream):
            MC.chains.init_log_like()
    done = 0
    while done < nmc:                                  # all problems advance chunk by chunk, concurrently
        n = min(chunk, nmc - done)
        for lag, MC, stream in runs:
            with torch.cuda.stream(stream):
                MC.chains.run(n, MC.mc_params())        # asynchronous launch on this problem's stream
        done += n
    outfiles = []
    for lag, MC, stream in runs:
        stream.synchronize()
        MC.steps_done = nmc
        MC.sync_from_device()
        n = MC.model.dim_v if nbins_tag is None else nbins_tag
        lagtag = ("%d" % lag) if float(lag).is_integer() else ("%g" % lag)
        out = os.path.join(outdir, "out.nbins%d.%s.%s.dat" % (n, lagtag, "pbc" if pbc else "nopbc"))
        with open(out, "w") as f:
            MC.print_laststate(f, final=True)
        np.savez_compressed(out + ".chains.npz", log_like=MC.chain_log_like, v=MC.chain_v, w=MC.chain_w,
                            naccv=MC.chain_naccv, naccw=MC.chain_naccw, status=MC.chain_status)
        outfiles.append(out)
    return outfiles

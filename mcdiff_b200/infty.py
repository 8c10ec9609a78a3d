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
    """One independent single-lag problem per entry of `files_by_lag` ({lag: filename}); writes
    out.nbins<N>.<lag>.pbc.dat (chain 0 of that problem) and .chains.npz per lag, in the naming mcdiff-infty
    parses.  Returns the list of output files (ascending lag) on every rank.

    One GPU: all problems advance concurrently, one CUDA stream each.  Under torchrun the
    len(files_by_lag) x nchains (problem, chain) pairs are block-partitioned over the ranks
    (parallel.pair_shard; SURVEY.md 8e, the reference runs them one after the other,
    examples/multipleruns.sh:74-91): a rank creates device chains only for the problems it owns a piece of,
    nothing is exchanged while sampling, and one gather per problem brings the per-chain rows to rank 0, which
    writes the files.  The Philox stream of a pair is keyed by (seed + problem, chain), so the files do not
    depend on the number of GPUs."""
    import torch
    from . import _capi as K
    from .MCState import MCState
    from .parallel import dist_info, gather_ranges_to_root, pair_shard
    from .transitions import Transitions

    rank, world, _ = dist_info()
    if rank == 0:
        os.makedirs(outdir, exist_ok=True)
    lags = sorted(files_by_lag)
    shards = [pair_shard(len(lags), nchains, r, world) for r in range(world)]      # [rank][problem] -> (lo, hi)
    runs = []
    for k, lag in enumerate(lags):
        MC = MCState(pbc, -1, nchains=nchains, device=device, seed=seed + k, rank=rank, world=world,
                     chain_range=shards[rank][k])
        MC.set_MC_params(dv, dw, 0.5, D0, 0.1, temp, nmc, nmc_update, False, -1, temp_end=temp_end)
        data = Transitions([files_by_lag[lag]])
        MC.set_model(model, data, ncosF, ncosD, 0, None)
        stream = None
        if MC.nchains > 0:
            dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
            stream = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(stream):                # uploads and launches of this problem share one stream
                MC._ensure_device()
        runs.append((lag, MC, stream))
    mine = [r for r in runs if r[2] is not None]
    for lag, MC, stream in mine:                           # first likelihoods, one stream each
        with torch.cuda.stream(stream):
            MC.chains.init_log_like()
    done = 0
    while done < nmc:                                      # all problems advance chunk by chunk, concurrently
        n = min(chunk, nmc - done)
        for lag, MC, stream in mine:
            with torch.cuda.stream(stream):
                MC.chains.run(n, MC.mc_params())            # asynchronous launch on this problem's stream
        done += n
    outfiles = []
    for k, (lag, MC, stream) in enumerate(runs):
        m = MC.model
        nF, nD = max(m.ncosF, 0), max(m.ncosD, 0)
        nv, nw = m.dim_v, m.dim_w
        if stream is not None:
            stream.synchronize()
            st = MC.chains.host_state()
            zc = np.zeros((MC.nchains, 0))
            local = np.concatenate([st["scal"], st["counters"].astype(np.float64), st["v"], st["w"],
                                    st["vcoef"] if nF > 0 else zc, st["wcoef"] if nD > 0 else zc,
                                    st["naccv_coeff"].astype(np.float64) if nF > 0 else zc,
                                    st["naccw_coeff"].astype(np.float64) if nD > 0 else zc], axis=1)
        else:
            local = np.zeros((0, 16 + nv + nw + 2 * nF + 2 * nD))
        rows = gather_ranges_to_root(local, [shards[r][k] for r in range(world)])
        n = nv if nbins_tag is None else nbins_tag
        lagtag = ("%d" % lag) if float(lag).is_integer() else ("%g" % lag)
        out = os.path.join(outdir, "out.nbins%d.%s.%s.dat" % (n, lagtag, "pbc" if pbc else "nopbc"))
        outfiles.append(out)
        if rank != 0:
            continue
        o = 16 + nv + nw
        MC.steps_done = nmc
        MC.set_host_state(dict(
            scal=rows[:, :8], counters=rows[:, 8:16].astype(np.int64), v=rows[:, 16:16 + nv],
            w=rows[:, 16 + nv:o], vcoef=rows[:, o:o + nF] if nF > 0 else None,
            wcoef=rows[:, o + nF:o + nF + nD] if nD > 0 else None,
            naccv_coeff=rows[:, o + nF + nD:o + 2 * nF + nD].astype(np.int64),
            naccw_coeff=rows[:, o + 2 * nF + nD:].astype(np.int64)))
        with open(out, "w") as f:
            MC.print_laststate(f, final=True)
        np.savez(out + ".chains.npz", log_like=MC.chain_log_like, v=MC.chain_v, w=MC.chain_w,
                            naccv=MC.chain_naccv, naccw=MC.chain_naccw, status=MC.chain_status)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()                                    # the files exist when any rank returns
    return outfiles


def main(argv=None):
    """`mcdiff-lagscan`: examples/multipleruns.sh + scripts/mcdiff-infty in one command -- one single-lag run per
    transition file (the lag time is read from the file header), all of them at once on the GPUs of the box
    (torchrun shards the (problem, chain) pairs), then the extrapolation D(lag -> infinity)."""
    from optparse import OptionParser
    from .transitions import Transitions
    p = OptionParser(usage="usage: %prog [options] SYSTEM transitions.lag1.dat [transitions.lag2.dat ...]")
    p.add_option("-o", "--outdir", dest="outdir", default=".")
    p.add_option("--nopbc", dest="pbc", default=True, action="store_false")
    p.add_option("-n", "--nmc", dest="nmc", default=1000, type="int")
    p.add_option("--nmc_update", dest="nmc_update", default=100, type="float")
    p.add_option("-T", dest="temp", default=1.0, type="float")
    p.add_option("--Tend", dest="temp_end", default=0.01, type="float")
    p.add_option("--dv", dest="dv", default=0.5, type="float")
    p.add_option("--dw", dest="dw", default=0.5, type="float")
    p.add_option("--D0", dest="D0", default=1.0, type="float")
    p.add_option("--model", dest="model", default="CosinusModel")
    p.add_option("--ncosF", dest="ncosF", default=10, type="int")
    p.add_option("--ncosD", dest="ncosD", default=6, type="int")
    p.add_option("-s", "--seed", dest="seed", default=1, type="int")
    p.add_option("--nchains", dest="nchains", default=256, type="int", help="chains per lag time")
    p.add_option("--dtc", dest="dtc", default=1.0, type="float", help="time unit of the lag in the file names")
    o, args = p.parse_args(argv)
    if len(args) < 3:
        p.error("SYSTEM and at least two transition files are needed")
    device = None
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        device = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    by_lag = {}
    for fn in args[1:]:
        by_lag[float(Transitions([fn]).list_lt[0])] = fn
    outs = run_lag_scan(by_lag, o.outdir, pbc=o.pbc, model=o.model, ncosF=o.ncosF, ncosD=o.ncosD, dv=o.dv, dw=o.dw,
                        D0=o.D0, temp=o.temp, temp_end=o.temp_end, nmc=o.nmc, nmc_update=o.nmc_update, seed=o.seed,
                        nchains=o.nchains, device=device)
    from .parallel import dist_info
    if dist_info()[0] == 0:
        lts, Dreal, t0 = extrapolate(outs, os.path.join(o.outdir, args[0] + "."), dtc=o.dtc)
        print("lag times", lts)
        print("Dreal[:10]", Dreal[:10])

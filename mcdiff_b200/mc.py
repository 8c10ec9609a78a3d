"""Driver: `find_parameters` and `do_mc_cycles` with the reference's signatures
(lib/mc.py:20-125), extended by the number of chains and GPUs.

    find_parameters(filenames, pbc, model, dv, dw, dwrad, D0, dtimezero, temp, temp_end, nmc,
                    nmc_update, seed, outfile, ncosF, ncosD, ncosDrad, move_timezero, initfile, k,
                    lmax, reduction, pull, nchains=1, ...)

Output files (formats of lib/log.py:288-376, consumed by outreading / plotresults /
mcdiff-infty / --initf):
    <outfile>               final state of chain 0 (the reference's single-chain file)
    <outfile>.pic           pickled Logger with the samples of ALL chains pooled
    <outfile>.pic.ave.dat   pooled posterior mean and standard deviation
    <outfile>.chains.npz    final state, log-likelihood and acceptance counters of every chain
"""
from __future__ import annotations

import sys

import numpy as np

from .MCState import MCState
from .log import Logger, write_average_from_pic
from .parallel import all_gather_rows, dist_info
from .transitions import RadTransitions, Transitions


def do_mc_cycles(MC, logger, chunk=1000):
    """The loop of lib/mc.py:20-51: all chains advance on the GPU in chunks of `chunk` steps
    (a multiple of 100 so that every Logger row lands at the end of a chunk boundary or inside
    one); after each chunk the sample rows are copied into `logger` and a progress line is
    printed for chain 0."""
    print("\n MC-move log-like acc(v) acc(w)")
    MC.init_log_like()
    MC.sync_from_device()
    m = MC.model
    if MC.do_radial:
        return _do_mc_cycles_radial(MC, logger, chunk)
    # row 0 = initial state (logger.log(0, MC), lib/mc.py:25)
    nv = m.ncosF if m.ncosF > 0 else m.dim_v
    nw = m.ncosD if m.ncosD > 0 else m.dim_w
    pv = MC.chain_v_coeff if m.ncosF > 0 else MC.chain_v
    pw = MC.chain_w_coeff if m.ncosD > 0 else MC.chain_w
    row0 = np.zeros((1, 6 + nv + nw))
    for c in range(MC.nchains):
        row0[0, :6] = (MC.chain_log_like[c], MC.chain_timezero[c], MC.chain_dv[c], MC.chain_dw[c], MC.dtimezero,
                       MC.chain_temp[c])
        row0[0, 6:6 + nv] = pv[c]
        row0[0, 6 + nv:] = pw[c]
        logger.fill_rows(c, 0, row0, m)
    done = 0
    next_row = 1
    while done < MC.nmc:
        n = min(chunk, MC.nmc - done)
        samples = MC.advance(n)
        done += n
        if samples is not None and samples.shape[1] > 0:
            for c in range(MC.nchains):
                logger.fill_rows(c, next_row, samples[c], m)
            next_row += samples.shape[1]
        MC.print_intermediate(done - 1)
    logger.dwrad[:] = MC.dwrad


def _finish_radial(MC, logger, outfile, nmc, rank):
    local = np.concatenate([MC.chain_log_like[:, None], MC.chain_dwrad[:, None],
                            MC.chain_naccwrad[:, None].astype(float), MC.chain_status[:, None].astype(float),
                            MC.chain_wrad], axis=1)
    allrows = all_gather_rows(local, MC.nchains_total)
    if rank != 0:
        return MC, logger
    if outfile is None:
        MC.print_statistics(f=sys.stdout)
        MC.print_laststate(sys.stdout, final=True)
        picfile = "mc.pic"
    else:
        with open(outfile, "w") as f:
            MC.print_statistics(f=sys.stdout)
            MC.print_laststate(f, final=True)
        picfile = outfile + ".pic"
    logger.model = MC.model
    logger.dump(picfile)
    write_average_from_pic(picfile, picfile + ".ave.dat")
    np.savez_compressed((outfile or "mc") + ".chains.npz", log_like=allrows[:, 0], dwrad=allrows[:, 1],
                        naccwrad=allrows[:, 2], status=allrows[:, 3], wrad=allrows[:, 4:],
                        edges=np.asarray(MC.model.edges), nmc=nmc, seed=MC.seed)
    return MC, logger


def _do_mc_cycles_radial(MC, logger, chunk):
    """lib/mc.py:29-30: only MCState.mcmove_diffusion_radial is called."""
    m = MC.model
    ncos = getattr(m, "ncosDrad", 0)
    pw = MC.chain_wrad_coeff if ncos > 0 else MC.chain_wrad
    nw = ncos if ncos > 0 else m.dim_wrad
    row0 = np.zeros((1, 3 + nw))
    for c in range(MC.nchains):
        row0[0, :3] = (MC.chain_log_like[c], MC.chain_dwrad[c], MC.chain_temp[c])
        row0[0, 3:] = pw[c][:nw]
        logger.fill_rows_radial(c, 0, row0, MC)
    done, next_row = 0, 1
    while done < MC.nmc:
        n = min(chunk, MC.nmc - done)
        samples = MC.advance(n)
        done += n
        if samples is not None and samples.shape[1] > 0:
            for c in range(MC.nchains):
                logger.fill_rows_radial(c, next_row, samples[c], MC)
            next_row += samples.shape[1]
        MC.print_intermediate(done - 1)


def find_parameters(filenames, pbc, model, dv, dw, dwrad, D0, dtimezero, temp, temp_end, nmc, nmc_update, seed,
                    outfile, ncosF, ncosD, ncosDrad, move_timezero, initfile, k, lmax, reduction, pull,
                    nchains=1, device=None, method="auto", jitter=0.0, chunk=1000):
    print("mcdiff_b200: free energy and diffusion profiles from transition counts, many chains on B200")
    rank, world, _ = dist_info()
    MC = MCState(pbc, lmax, nchains=nchains, device=device, seed=seed, rank=rank, world=world, method=method,
                 jitter=jitter)
    MC.set_MC_params(dv, dw, dwrad, D0, dtimezero, temp, nmc, nmc_update, move_timezero, k, temp_end=temp_end)
    data = RadTransitions(filenames) if MC.do_radial else Transitions(filenames, reduction=reduction)
    MC.set_model(model, data, ncosF, ncosD, ncosDrad, pull)
    if initfile is not None:
        MC.use_initfile(initfile)
        MC.print_MC_params(sys.stdout)
        MC.print_coeffs_laststate(sys.stdout)
    logger = Logger(MC, nchains=MC.nchains)
    do_mc_cycles(MC, logger, chunk=chunk)

    # ---- one collective: per-chain summaries of every rank (lib/mc.py has a single chain) ----
    m = MC.model
    if MC.do_radial:
        return _finish_radial(MC, logger, outfile, nmc, rank)
    pv = MC.chain_v_coeff if m.ncosF > 0 else MC.chain_v
    pw = MC.chain_w_coeff if m.ncosD > 0 else MC.chain_w
    local = np.concatenate([MC.chain_log_like[:, None], MC.chain_dv[:, None], MC.chain_dw[:, None],
                            MC.chain_naccv[:, None].astype(float), MC.chain_naccw[:, None].astype(float),
                            MC.chain_nacctimezero[:, None].astype(float), MC.chain_status[:, None].astype(float),
                            MC.chain_v, MC.chain_w, pv, pw], axis=1)
    allrows = all_gather_rows(local, MC.nchains_total)
    if rank != 0:
        return MC, logger

    def write_final(f):
        MC.print_statistics(f=sys.stdout)
        MC.print_laststate(f, final=True)

    if outfile is None:
        write_final(sys.stdout)
        picfile = "mc.pic"
    else:
        with open(outfile, "w") as f:
            write_final(f)
        picfile = outfile + ".pic"
    logger.model = MC.model
    logger.dump(picfile)
    write_average_from_pic(picfile, picfile + ".ave.dat")
    nv, nw = m.dim_v, m.dim_w
    np.savez_compressed((outfile or "mc") + ".chains.npz",
                        log_like=allrows[:, 0], dv=allrows[:, 1], dw=allrows[:, 2], naccv=allrows[:, 3],
                        naccw=allrows[:, 4], nacctimezero=allrows[:, 5], status=allrows[:, 6],
                        v=allrows[:, 7:7 + nv], w=allrows[:, 7 + nv:7 + nv + nw],
                        v_param=allrows[:, 7 + nv + nw:7 + nv + nw + pv.shape[1]],
                        w_param=allrows[:, 7 + nv + nw + pv.shape[1]:], edges=np.asarray(m.edges), nmc=nmc,
                        seed=MC.seed)
    return MC, logger

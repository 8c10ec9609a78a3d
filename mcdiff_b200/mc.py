"""Driver: `find_parameters` and `do_mc_cycles` with the reference's signatures
(lib/mc.py:20-125), extended by the number of chains and GPUs.

    find_parameters(filenames, pbc, model, dv, dw, dwrad, D0, dtimezero, temp, temp_end, nmc,
                    nmc_update, seed, outfile, ncosF, ncosD, ncosDrad, move_timezero, initfile, k,
                    lmax, reduction, pull, nchains=1, ...)

Output files (formats of lib/log.py:288-376, consumed by outreading / plotresults /
mcdiff-infty / --initf):
    <outfile>               final state of chain 0 (the reference's single-chain file)
    <outfile>.pic           pickled Logger with the samples of ALL chains of ALL ranks pooled (global chain order)
    <outfile>.pic.ave.dat   pooled posterior mean and standard deviation
    <outfile>.chains.npz    final state, log-likelihood, acceptance and step counters of every chain
                            (`--resume` continues every chain exactly from it)

Under torchrun the chains shard over the ranks (parallel.shard_range); nothing is exchanged while sampling.
Afterwards two collectives run: an all_gather of the per-chain state/acceptance rows and a gather of the Logger
rows on rank 0, which writes the files -- so 1 GPU and N GPUs with the same --nchains give identical files.
"""
from __future__ import annotations

import sys

import numpy as np

from .MCState import MCState
from .log import Logger, write_average_from_pic
from .parallel import all_gather_rows, dist_info, gather_rows_to_root
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
    row0 = np.zeros((MC.nchains, 1, 6 + nv + nw))
    row0[:, 0, 0], row0[:, 0, 1] = MC.chain_log_like, MC.chain_timezero
    row0[:, 0, 2], row0[:, 0, 3] = MC.chain_dv, MC.chain_dw
    row0[:, 0, 4], row0[:, 0, 5] = MC.dtimezero, MC.chain_temp
    row0[:, 0, 6:6 + nv] = pv
    row0[:, 0, 6 + nv:] = pw
    logger.fill_rows_all(0, row0, m)
    done = 0
    next_row = 1
    while done < MC.nmc:
        n = min(chunk, MC.nmc - done)
        samples = MC.advance(n)
        done += n
        if samples is not None and samples.shape[1] > 0:
            logger.fill_rows_all(next_row, samples, m)
            next_row += samples.shape[1]
        MC.print_intermediate(done - 1)
    logger.dwrad[:] = MC.dwrad


def _pool_logger(MC, logger, rank):
    """Rows of every chain of every rank -> one Logger on rank 0 (global chain order), so that `.pic` and
    `.pic.ave.dat` are the pooled posterior of ALL chains whatever the number of GPUs
    (lib/log.py:90-106, lib/outreading.py:281-301 take their statistics over the rows of this object)."""
    rows = gather_rows_to_root(logger.pack_rows(), MC.nchains_total)
    if rank != 0:
        return None
    if MC.nchains_total == logger.nchains:
        return logger
    pooled = Logger(MC, nchains=MC.nchains_total)
    pooled.unpack_rows(rows)
    return pooled


def _write_outputs(MC, pooled, outfile):
    """Rank 0: the reference's result file for chain 0, the pickled pooled Logger and its average."""
    if outfile is None:
        MC.print_statistics(f=sys.stdout)
        MC.print_laststate(sys.stdout, final=True)
        picfile = "mc.pic"
    else:
        with open(outfile, "w") as f:
            MC.print_statistics(f=sys.stdout)
            MC.print_laststate(f, final=True)
        picfile = outfile + ".pic"
    pooled.model = MC.model
    pooled.dump(picfile)
    # the reference re-reads the pickle it has just written (write_average_from_pic, lib/log.py:380-387); the Logger in
    # memory holds the same rows, and un-pickling 8192 chains costs more than the statistics themselves
    with open(picfile + ".ave.dat", "w") as f:
        pooled.print_average(pooled.model, f=f)


def _finish_radial(MC, logger, outfile, nmc, rank):
    ncos = getattr(MC.model, "ncosDrad", 0)
    local = np.concatenate([MC.chain_scal, MC.chain_counters.astype(np.float64), MC.chain_wrad,
                            MC.chain_wrad_coeff if ncos > 0 else np.zeros((MC.nchains, 0)),
                            MC.chain_naccwrad_coeff.astype(np.float64) if ncos > 0 else np.zeros((MC.nchains, 0))],
                           axis=1)
    allrows = all_gather_rows(local, MC.nchains_total)
    pooled = _pool_logger(MC, logger, rank)
    MC.set_global_summary(log_like=allrows[:, 0], acc=allrows[:, 8] / float(max(1, nmc)))
    if rank != 0:
        return MC, logger
    _write_outputs(MC, pooled, outfile)
    n = MC.model.dim_wrad
    np.savez((outfile or "mc") + ".chains.npz", log_like=allrows[:, 0], dwrad=allrows[:, 1],
                        naccwrad=allrows[:, 8], status=allrows[:, 11], scal=allrows[:, :8],
                        counters=allrows[:, 8:16].astype(np.int64), wrad=allrows[:, 16:16 + n],
                        wrad_coeff=allrows[:, 16 + n:16 + n + ncos],
                        naccwrad_coeff=allrows[:, 16 + n + ncos:].astype(np.int64),
                        edges=np.asarray(MC.model.edges), nmc=nmc, seed=MC.seed, steps_done=MC.steps_done, radial=1)
    return MC, pooled


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
                    nchains=1, device=None, method="auto", jitter=0.0, chunk=1000, resume=None):
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
    if resume is not None:
        MC.use_chainfile(resume)
    logger = Logger(MC, nchains=MC.nchains)
    do_mc_cycles(MC, logger, chunk=chunk)

    # ---- the collectives of the design, after sampling: per-chain state + acceptance counters
    #      on every rank, the Logger rows of all chains on rank 0 (lib/mc.py has a single chain) ----
    m = MC.model
    if MC.do_radial:
        return _finish_radial(MC, logger, outfile, nmc, rank)
    nF, nD = max(m.ncosF, 0), max(m.ncosD, 0)
    zc = np.zeros((MC.nchains, 0))
    local = np.concatenate([MC.chain_scal, MC.chain_counters.astype(np.float64), MC.chain_v, MC.chain_w,
                            MC.chain_v_coeff if nF > 0 else zc, MC.chain_w_coeff if nD > 0 else zc,
                            MC.chain_naccv_coeff.astype(np.float64) if nF > 0 else zc,
                            MC.chain_naccw_coeff.astype(np.float64) if nD > 0 else zc], axis=1)
    allrows = all_gather_rows(local, MC.nchains_total)
    pooled = _pool_logger(MC, logger, rank)
    from . import _capi as K
    MC.set_global_summary(log_like=allrows[:, K.S_LL],
                          acc=(allrows[:, 8 + K.C_NACCV] + allrows[:, 8 + K.C_NACCW]) / float(max(1, nmc)))
    if rank != 0:
        return MC, logger
    _write_outputs(MC, pooled, outfile)
    nv, nw = m.dim_v, m.dim_w
    o = 16
    v, w = allrows[:, o:o + nv], allrows[:, o + nv:o + nv + nw]
    o += nv + nw
    vc, wc = allrows[:, o:o + nF], allrows[:, o + nF:o + nF + nD]
    o += nF + nD
    np.savez((outfile or "mc") + ".chains.npz",
                        log_like=allrows[:, K.S_LL], dv=allrows[:, K.S_DV], dw=allrows[:, K.S_DW],
                        naccv=allrows[:, 8 + K.C_NACCV], naccw=allrows[:, 8 + K.C_NACCW],
                        nacctimezero=allrows[:, 8 + K.C_NACCT0], status=allrows[:, 8 + K.C_STATUS],
                        v=v, w=w, v_param=vc if nF > 0 else v, w_param=wc if nD > 0 else w,
                        # exact resume: the whole per-chain state; the Philox counter of a chain is
                        # (step, global chain index), so `counters[:, C_STEP]` + `seed` continue every stream
                        scal=allrows[:, :8], counters=allrows[:, 8:16].astype(np.int64), v_coeff=vc, w_coeff=wc,
                        naccv_coeff=allrows[:, o:o + nF].astype(np.int64),
                        naccw_coeff=allrows[:, o + nF:o + nF + nD].astype(np.int64),
                        edges=np.asarray(m.edges), nmc=nmc, seed=MC.seed, steps_done=MC.steps_done, radial=0)
    return MC, pooled

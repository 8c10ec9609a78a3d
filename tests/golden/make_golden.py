#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (imported read-only from /root/reference/lib as package ``mcdiff``).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Outputs (committed):
    tests/golden/loglike_golden.npz   log-likelihood / propagator known answers
    tests/golden/replay_golden.npz    tape-driven MC trajectories of the reference
    tests/golden/radial_golden.npz    radial (Bessel) likelihood known answers
    tests/golden/radial_replay_golden.npz  tape-driven radial MC trajectories of the reference

Nothing here is product code and nothing is copied from the reference: the
reference functions are *called* and their outputs recorded.
"""
import contextlib
import io
import zlib
import os
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    shim = tempfile.mkdtemp(prefix="mcdiff_ref_")
    os.symlink(os.path.join(REF, "lib"), os.path.join(shim, "mcdiff"))
    sys.path.insert(0, shim)
    import warnings
    warnings.simplefilter("ignore")
    import mcdiff  # noqa: F401
    return shim


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


class TapeRNG:
    """Feeds np.random.rand/randint/random calls of the reference from a tape.

    tape_u[s] = (u_choice, u_delta, u_accept, u_t0_delta, u_t0_accept); which
    column a ``random()`` call receives is decided by the calling reference
    function (mcmove_timezero vs mcmove_potential/diffusion/diffusion_radial).
    """

    def __init__(self, tape_u, tape_idx, radial=False):
        self.u = tape_u
        self.idx = tape_idx
        self.s = -1
        self.calls_move = 0
        self.calls_t0 = 0
        self.radial = radial

    def _next_step(self):
        self.s += 1
        self.calls_move = 0
        self.calls_t0 = 0

    def rand(self):
        self._next_step()
        return self.u[self.s, 0]

    def randint(self, lo, hi):
        if self.radial:
            self._next_step()
        v = int(self.idx[self.s])
        assert lo <= v < hi, (lo, v, hi)
        return v

    def random(self):
        caller = sys._getframe(1).f_code.co_name
        if caller == "mcmove_timezero":
            col = 3 + self.calls_t0
            self.calls_t0 += 1
        else:
            col = 1 + self.calls_move
            self.calls_move += 1
        return self.u[self.s, col]


def run_reference_replay(files, *, pbc, ncosF, ncosD, dv, dw, temp, temp_end, nmc,
                         nmc_update, move_timezero, dtimezero, k, tape_u, tape_idx,
                         initfile=None, D0=1.0, pull=None):
    import mcdiff.MCState as mcs
    import mcdiff.mc as mcmod
    from mcdiff.transitions import Transitions
    from mcdiff.log import Logger

    rec = dict(move=[], accepted=[], accepted_t0=[], ll_after=[], widths=[])
    with quiet():
        MC = mcs.MCState(pbc, -1)
        MC.set_MC_params(dv, dw, 0.5, D0, dtimezero, temp, nmc, nmc_update, move_timezero, k,
                         temp_end=temp_end)
        data = Transitions(files)
        MC.set_model("CosinusModel", data, ncosF, ncosD, 0, pull)
        if initfile is not None:
            MC.use_initfile(initfile)
        logger = Logger(MC)
        v0 = MC.model.v.copy()
        w0 = MC.model.w.copy()
        vc0 = MC.model.v_coeff.copy() if ncosF > 0 else np.zeros(0)
        wc0 = MC.model.w_coeff.copy() if ncosD > 0 else np.zeros(0)

        rng = TapeRNG(tape_u, tape_idx)
        saved = (np.random.rand, np.random.randint, np.random.random)
        np.random.rand, np.random.randint, np.random.random = rng.rand, rng.randint, rng.random
        try:
            # the body of do_mc_cycles (lib/mc.py:20-51), instrumented between calls only
            MC.init_log_like()
            ll0 = float(MC.log_like)
            logger.log(0, MC)
            for imc in range(MC.nmc):
                nv, nw, nt = MC.naccv, MC.naccw, MC.nacctimezero
                choice = np.random.rand()
                if choice < 0.5 and MC.model.ncosF != 1 and MC.dv > 0.:
                    MC.mcmove_potential()
                    rec["move"].append(1)
                elif MC.dw > 0.:
                    MC.mcmove_diffusion()
                    rec["move"].append(2)
                else:
                    rec["move"].append(0)
                if MC.move_timezero and MC.dtimezero > 0:
                    MC.mcmove_timezero()
                rec["accepted"].append(int(MC.naccv + MC.naccw - nv - nw))
                rec["accepted_t0"].append(int(MC.nacctimezero - nt))
                rec["ll_after"].append(float(MC.log_like))
                logger.log(imc + 1, MC)
                dv_b, dw_b = MC.dv, MC.dw
                MC.update_movewidth(imc)
                MC.update_temp(imc)
                if (MC.dv, MC.dw) != (dv_b, dw_b):
                    rec["widths"].append((imc, MC.dv, MC.dw))
        finally:
            np.random.rand, np.random.randint, np.random.random = saved
    out = dict(
        ll0=ll0, v0=v0, w0=w0, vc0=vc0, wc0=wc0,
        move=np.array(rec["move"], dtype=np.int8),
        accepted=np.array(rec["accepted"], dtype=np.int8),
        accepted_t0=np.array(rec["accepted_t0"], dtype=np.int8),
        ll_after=np.array(rec["ll_after"]),
        widths=np.array(rec["widths"], dtype=np.float64).reshape(-1, 3),
        v=MC.model.v.copy(), w=MC.model.w.copy(),
        vc=MC.model.v_coeff.copy() if ncosF > 0 else np.zeros(0),
        wc=MC.model.w_coeff.copy() if ncosD > 0 else np.zeros(0),
        timezero=float(MC.model.timezero), ll=float(MC.log_like),
        dv=float(MC.dv), dw=float(MC.dw), temp=float(MC.temp),
        naccv=int(MC.naccv), naccw=int(MC.naccw), nacct0=int(MC.nacctimezero),
        naccv_coeff=np.array(MC.naccv_coeff) if ncosF > 0 else np.zeros(0, dtype=np.int64),
        naccw_coeff=np.array(MC.naccw_coeff) if ncosD > 0 else np.zeros(0, dtype=np.int64),
        log_ll=logger.log_like.copy(), log_dv=logger.dv.copy(), log_dw=logger.dw.copy(),
        log_timezero=logger.timezero.copy(),
        log_pv=(logger.v_coeff if ncosF > 0 else logger.v).copy(),
        log_pw=(logger.w_coeff if ncosD > 0 else logger.w).copy(),
        wunit=float(MC.model.wunit), edges=np.array(MC.model.edges),
        counts=np.array(data.list_trans).astype(np.int32), lagtimes=np.array(data.list_lt, dtype=np.float64),
    )
    if pull is not None:
        out["dF"] = float(MC.pull)     # dF per bin the reference derived from --pull (lib/MCState.py:65-67)
    return out


def tape_for(rng, nsteps, n, dim_w, ncosF, ncosD, dv, dw):
    u = rng.random((nsteps, 5))
    idx = np.zeros(nsteps, dtype=np.int32)
    for s in range(nsteps):
        if u[s, 0] < 0.5 and ncosF != 1 and dv > 0:
            idx[s] = rng.integers(0, n) if ncosF <= 0 else rng.integers(1, ncosF)
        else:
            idx[s] = rng.integers(0, dim_w) if ncosD <= 0 else rng.integers(0, ncosD)
    return u, idx


def main():
    _import_reference()
    from mcdiff.utils import log_like_lag, init_rate_matrix
    from mcdiff.transitions import Transitions, cut_transitions_square
    from mcdiff.outreading import read_Fcoeffs, read_Dcoeffs
    from mcdiff.twod import rad_log_like_lag, setup_bessel_functions, propagator_radial_diffusion
    import scipy.linalg

    ex = os.path.join(REF, "examples")
    fA = os.path.join(ex, "data/HexdWat/transitions.nbins100.20.pbc.A.dat")
    fM = os.path.join(ex, "manipulate-transition-matrices/transitions.merge.dat")
    pbt = [os.path.join(ex, "create-transition-matrix/pbctrans/transitions.nbins100.%d.pbc.dat" % l)
           for l in (10, 20, 30, 40)]
    datA = os.path.join(ex, "output/HexdWat/A/out.ncosf10d6.run1/out.nbins100.20.pbc.dat")
    datM = os.path.join(ex, "output/HexdWat/out.ncosf10d6.run1/out.nbins100.20.pbc.dat")

    def cos_profiles(n, dim_w, vcoef, wcoef):
        x = np.arange(n)
        vb = np.array([np.cos(2 * k * np.pi * (x + 0.5) / n) / (k + 1) for k in range(len(vcoef))]).T
        xw = np.arange(dim_w)
        wb = np.array([np.cos(2 * k * np.pi * (xw + 1.0) / n) / (k + 1) for k in range(len(wcoef))]).T
        return vb @ vcoef, wb @ wcoef

    G = {}
    with quiet():
        trA = Transitions([fA])
        trM = Transitions([fM])
        trP = Transitions(pbt)
    n = 100
    dz = trA.edges[1] - trA.edges[0]
    wunit = np.log(dz ** 2 / 1.0)
    G["A_counts"] = trA.list_trans.astype(np.int32)
    G["A_lags"] = trA.list_lt.astype(np.float64)
    G["A_edges"] = trA.edges
    G["M_counts"] = trM.list_trans.astype(np.int32)
    G["M_lags"] = trM.list_lt.astype(np.float64)
    G["P_counts"] = trP.list_trans.astype(np.int32)
    G["P_lags"] = trP.list_lt.astype(np.float64)
    G["P_edges"] = trP.edges

    # ---- K1..K4 (archived log values) ----
    flat_v = np.zeros(n)
    flat_w = np.zeros(n) + (np.log(1.0) - wunit)
    vcA = read_Fcoeffs(datA, final=True)
    wcA = read_Dcoeffs(datA, final=True)
    wcA[0] -= wunit
    vA, wA = cos_profiles(n, n, vcA, wcA)
    dzM = trM.edges[1] - trM.edges[0]
    wunitM = np.log(dzM ** 2)
    vcM = read_Fcoeffs(datM, final=True)
    wcM = read_Dcoeffs(datM, final=True)
    wcM[0] -= wunitM
    vM, wM = cos_profiles(n, n, vcM, wcM)
    G["K_archived"] = np.array([-3643153.78029, -3480590.26443, -14577042.6175, -14048561.0464])
    G["K_v"] = np.array([flat_v, vA, flat_v, vM])
    G["K_w"] = np.array([flat_w, wA, np.zeros(n) - wunitM, wM])
    G["K_ll"] = np.array([
        log_like_lag(n, 1, flat_v, flat_w, trA.list_lt, trA.list_trans, True, None),
        log_like_lag(n, 1, vA, wA, trA.list_lt, trA.list_trans, True, None),
        log_like_lag(n, 1, flat_v, np.zeros(n) - wunitM, trM.list_lt, trM.list_trans, True, None),
        log_like_lag(n, 1, vM, wM, trM.list_lt, trM.list_trans, True, None)])
    G["K2_P"] = scipy.linalg.expm(20.0 * init_rate_matrix(n, vA, wA, True))
    G["wunit_A"] = np.float64(wunit)

    # ---- pbctrans 4 lags, flat ----
    dzP = trP.edges[1] - trP.edges[0]
    wP = np.zeros(n) - np.log(dzP ** 2)
    G["P_flat_w"] = wP
    G["P_flat_ll"] = np.float64(log_like_lag(n, 4, flat_v, wP, trP.list_lt, trP.list_trans, True, None))

    # ---- random profiles on HexdWat A (pbc, 1 lag) and pbctrans (4 lags) ----
    rng = np.random.default_rng(0)
    nb = 24
    rv = rng.normal(0.0, 1.0, (nb, n))
    rw = flat_w[None, :] + rng.normal(0.0, 0.3, (nb, n))
    G["rand_v"] = rv
    G["rand_w"] = rw
    G["rand_ll_A"] = np.array([log_like_lag(n, 1, rv[b], rw[b], trA.list_lt, trA.list_trans, True, None)
                               for b in range(nb)])
    rwP = wP[None, :] + rng.normal(0.0, 0.3, (nb, n))
    G["rand_wP"] = rwP
    G["rand_ll_P"] = np.array([log_like_lag(n, 4, rv[b], rwP[b], trP.list_lt, trP.list_trans, True, None)
                               for b in range(nb)])
    G["rand_P0"] = scipy.linalg.expm(20.0 * init_rate_matrix(n, rv[0], rw[0], True))

    # ---- nopbc: cut of HexdWat A written with the reference's own cutter ----
    tmp = tempfile.mkdtemp(prefix="mcd_golden_")
    fcut = os.path.join(tmp, "cut.dat")
    with quiet():
        cut_transitions_square(fA, 20, 80, fcut)
        trC = Transitions([fcut])
    nc = trC.dim_trans
    G["C_counts"] = trC.list_trans.astype(np.int32)
    G["C_lags"] = trC.list_lt.astype(np.float64)
    G["C_edges"] = trC.edges
    cv = rng.normal(0.0, 1.0, (8, nc))
    cw = (np.log(1.0) - np.log((trC.edges[1] - trC.edges[0]) ** 2)) + rng.normal(0.0, 0.3, (8, nc - 1))
    G["C_v"] = cv
    G["C_w"] = cw
    G["C_ll"] = np.array([log_like_lag(nc, 1, cv[b], cw[b], trC.list_lt, trC.list_trans, False, None)
                          for b in range(8)])
    G["C_P0"] = scipy.linalg.expm(20.0 * init_rate_matrix(nc, cv[0], cw[0], False))

    # ---- small odd-sized pbc problem (n = 37) from a reduced HexdWat block, 2 lags ----
    ns = 37
    sub = trA.list_trans[0][30:30 + ns, 30:30 + ns]
    Scnt = np.array([sub, sub[::-1, ::-1]]).astype(np.int32)
    Slags = np.array([5.0, 12.5])
    sv = rng.normal(0.0, 0.7, (6, ns))
    sw = 1.2 + rng.normal(0.0, 0.3, (6, ns))
    G["S_counts"] = Scnt
    G["S_lags"] = Slags
    G["S_v"] = sv
    G["S_w"] = sw
    G["S_ll"] = np.array([log_like_lag(ns, 2, sv[b], sw[b], Slags, Scnt, True, None) for b in range(6)])
    np.savez_compressed(os.path.join(HERE, "loglike_golden.npz"), **G)
    print("K values:", G["K_ll"], "vs archived", G["K_archived"])
    print("pbctrans flat:", G["P_flat_ll"])

    # ---- replay fixtures ----
    Rp = {}

    def add(tag, files, **kw):
        nmc = kw["nmc"]
        with quiet():
            tr = Transitions(files)
        nn = tr.dim_trans
        dim_w = nn if kw["pbc"] else nn - 1
        r = np.random.default_rng(zlib.crc32(tag.encode()))
        tu, ti = tape_for(r, nmc, nn, dim_w, kw["ncosF"], kw["ncosD"], kw["dv"], kw["dw"])
        out = run_reference_replay(files, tape_u=tu, tape_idx=ti, **kw)
        Rp[tag + "/tape_u"] = tu
        Rp[tag + "/tape_idx"] = ti
        for key, val in out.items():
            Rp[tag + "/" + key] = val
        for key in ("pbc", "ncosF", "ncosD", "dv", "dw", "temp", "nmc", "nmc_update", "move_timezero",
                    "dtimezero", "k"):
            Rp[tag + "/cfg_" + key] = np.float64(kw[key])
        Rp[tag + "/cfg_temp_end"] = np.float64(np.nan if kw["temp_end"] is None else kw["temp_end"])
        print(tag, "ll0", out["ll0"], "ll", out["ll"], "acc", out["naccv"], out["naccw"], out["nacct0"])

    base = dict(pbc=True, temp=1.0, temp_end=1.0, nmc_update=100.0, move_timezero=False,
                dtimezero=0.1, k=-1.0)
    add("cosA", [fA], ncosF=10, ncosD=6, dv=0.05, dw=0.05, nmc=400, **base)
    add("binP4", pbt, ncosF=0, ncosD=0, dv=0.5, dw=0.5, nmc=400, **base)
    b2 = dict(base)
    b2.update(move_timezero=True, dtimezero=0.5, temp_end=0.5)
    add("cosP4t0", pbt, ncosF=6, ncosD=4, dv=0.3, dw=0.3, nmc=300, **b2)
    b3 = dict(base)
    b3.update(pbc=False)
    add("binCut", [fcut], ncosF=0, ncosD=0, dv=0.1, dw=0.1, nmc=300, **b3)
    b4 = dict(base)
    b4.update(k=5.0)
    add("springP4", pbt, ncosF=0, ncosD=0, dv=0.3, dw=0.3, nmc=300, **b4)
    np.savez_compressed(os.path.join(HERE, "replay_golden.npz"), **Rp)

    # ---- radial fixture (synthetic cube; no radial data ships with the reference) ----
    Rd = {}
    nr, dim_rad, lmax = 16, 12, 10
    rng = np.random.default_rng(5)
    i = np.arange(nr)
    vr = 0.8 * np.cos(2 * np.pi * (i + 0.5) / nr)
    wr = np.zeros(nr) + np.log(1.0 / 0.5 ** 2)
    rate = init_rate_matrix(nr, vr, wr, True)
    zeros, table = setup_bessel_functions(lmax, dim_rad)
    wrad_true = np.zeros(nr) + np.log(0.6 / 0.5 ** 2)
    rlags = np.array([2.0, 5.0])
    cube = np.zeros((2, dim_rad, nr, nr))
    for il, lag in enumerate(rlags):
        P = propagator_radial_diffusion(nr, dim_rad, rate, wrad_true, lag, lmax, zeros, table)
        cube[il] = rng.poisson(2000.0 * np.maximum(P, 0) * np.exp(-vr)[None, None, :])
    wrads = wrad_true[None, :] + rng.normal(0, 0.3, (6, nr))
    Rd["rate"] = rate
    Rd["v"] = vr
    Rd["w"] = wr
    Rd["zeros"] = zeros
    Rd["table"] = table
    Rd["lags"] = rlags
    Rd["counts"] = cube
    Rd["wrad"] = wrads
    Rd["ll"] = np.array([rad_log_like_lag(nr, dim_rad, 2, rate, wrads[b], rlags, cube, None, lmax, zeros, table, 0.)
                         for b in range(6)])
    Rd["P0"] = propagator_radial_diffusion(nr, dim_rad, rate, wrads[0], rlags[0], lmax, zeros, table)
    np.savez_compressed(os.path.join(HERE, "radial_golden.npz"), **Rd)
    print("radial ll:", Rd["ll"])

    # ---- radial Monte-Carlo trajectories of the reference (lib/mc.py:29-30, lib/MCState.py:306-340) ----
    from mcdiff.tools.extract import write_Tmat_cube
    import mcdiff.MCState as mcs
    from mcdiff.transitions import RadTransitions
    from mcdiff.log import Logger
    edges = np.linspace(-4.0, 4.0, nr + 1)
    redges = np.arange(dim_rad) * 0.5
    rfiles = []
    for il, lag in enumerate(rlags):
        fn = os.path.join(tmp, "rad.%d.dat" % il)
        write_Tmat_cube(cube[il].astype(int), fn, float(lag), "pbc", edges=edges, redges=redges, dt=1.0, dn=int(lag))
        rfiles.append(fn)
    RR = {}
    for tag, ncosDrad, nmc in (("radbin", 0, 120), ("radcos", 4, 120)):
        r = np.random.default_rng(zlib.crc32(tag.encode()))
        tu = r.random((nmc, 5))
        ti = r.integers(0, nr if ncosDrad == 0 else ncosDrad, nmc).astype(np.int32)
        with quiet():
            MC = mcs.MCState(True, lmax)
            MC.set_MC_params(0.5, 0.5, 0.4, 1.0, 0.1, 1.0, nmc, 40.0, False, -1.0, temp_end=1.0)
            data = RadTransitions(rfiles)
            MC.set_model("RadCosinusModel", data, 0, 0, ncosDrad, None)
            logger = Logger(MC)
            wrad0 = MC.model.wrad.copy()
            wc0 = MC.model.wrad_coeff.copy() if ncosDrad > 0 else np.zeros(0)
            rng_ = TapeRNG(tu, ti, radial=True)
            saved = (np.random.rand, np.random.randint, np.random.random)
            np.random.rand, np.random.randint, np.random.random = rng_.rand, rng_.randint, rng_.random
            acc, ll_after, widths = [], [], []
            try:
                MC.init_log_like()
                ll0 = float(MC.log_like)
                logger.log(0, MC)
                for imc in range(nmc):
                    na = MC.naccwrad
                    MC.mcmove_diffusion_radial()
                    acc.append(int(MC.naccwrad - na))
                    ll_after.append(float(MC.log_like))
                    logger.log(imc + 1, MC)
                    d_b = MC.dwrad
                    MC.update_movewidth(imc)
                    MC.update_temp(imc)
                    if MC.dwrad != d_b:
                        widths.append((imc, MC.dwrad))
            finally:
                np.random.rand, np.random.randint, np.random.random = saved
        out = dict(tape_u=tu, tape_idx=ti, ll0=ll0, accepted=np.array(acc, dtype=np.int8), ll_after=np.array(ll_after),
                   widths=np.array(widths).reshape(-1, 2), wrad0=wrad0, wc0=wc0, wrad=np.array(MC.model.wrad),
                   wc=np.array(MC.model.wrad_coeff) if ncosDrad > 0 else np.zeros(0), ll=float(MC.log_like),
                   dwrad=float(MC.dwrad), naccwrad=int(MC.naccwrad), v=np.array(MC.model.v), w=np.array(MC.model.w),
                   zeros=np.array(MC.model.bessel0_zeros), table=np.array(MC.model.bessels),
                   counts=np.array(data.list_trans), lags=np.array(data.list_lt, dtype=np.float64),
                   wradunit=float(MC.model.wradunit), log_ll=logger.log_like.copy(),
                   log_wrad=(logger.wrad_coeff if ncosDrad > 0 else logger.wrad).copy(), dim_rad=dim_rad, lmax=lmax)
        for key, val in out.items():
            RR[tag + "/" + key] = val
        print(tag, "ll0", ll0, "ll", out["ll"], "acc", out["naccwrad"], "dwrad", out["dwrad"])
    np.savez_compressed(os.path.join(HERE, "radial_replay_golden.npz"), **RR)
    for f in ("loglike_golden.npz", "replay_golden.npz", "radial_golden.npz", "radial_replay_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()

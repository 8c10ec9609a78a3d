"""Monte-Carlo state of MANY chains, with the reference's MCState interface.

The reference's MCState (lib/MCState.py:23-491) holds one chain and evaluates one move per
Python call.  Here the same object fronts `nchains` independent chains that live on the GPU:
`set_MC_params`, `set_model`, `use_initfile`, `init_log_like` keep their names, arguments and
error behaviour; the per-move methods are replaced by `advance(nsteps)`, which runs the fused
CUDA loop (mcd_mc_run) for all chains at once.  Scalar attributes (`dv`, `dw`, `temp`,
`log_like`, `naccv`, ...) always mirror chain 0, so code written against the reference keeps
working; the `chain_*` arrays expose every chain.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import _capi as K
from .model import MODELS_1D, Model, RadCosinusModel, RadModel
from .outreading import read_Dcoeffs, read_Dradcoeffs, read_Drad, read_Fcoeffs, read_F_D_edges, read_dv_dw
from .parallel import shard_range
from .engine import to_host


def _string_vecs(n, pbc):
    """Eigenvectors of the 1-D Laplacian: directions of the smoothed moves (lib/utils.py:294-305)."""
    M = 2.0 * np.eye(n)
    i = np.arange(n - 1)
    M[i, i + 1] = -1.0
    M[i + 1, i] = -1.0
    if pbc:
        M[0, n - 1] = M[n - 1, 0] = -1.0
    else:
        M[0, 0] = M[n - 1, n - 1] = 1.0
    return np.linalg.eigh(M)[1]


class MCState:
    def __init__(self, pbc, lmax=-1, nchains=1, device=None, seed=None, rank=0, world=1, method="auto",
                 jitter=0.0, chain_range=None):
        self.pbc = pbc
        self.lmax = lmax
        self.do_radial = lmax > 0
        self.nchains_total = int(nchains)
        self.rank, self.world = int(rank), int(world)
        # chain_range: an explicit [lo, hi) of the global chain index (lag scans shard (problem, chain) pairs,
        # parallel.pair_shard), else the block partition of the chains over the ranks
        self.chain_lo, self.chain_hi = (shard_range(self.nchains_total, self.rank, self.world)
                                        if chain_range is None else (int(chain_range[0]), int(chain_range[1])))
        self.nchains = self.chain_hi - self.chain_lo
        self.device = device
        self.seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self.method = {"auto": K.METHOD_AUTO, "eigen": K.METHOD_EIGEN, "squaring": K.METHOD_SQUARING}[method]
        self.jitter = float(jitter)
        self.engine = None
        self.problem = None
        self.chains = None
        self.steps_done = 0

    # ---- lib/MCState.py:29-61 ------------------------------------------------------------
    def set_MC_params(self, dv, dw, dwrad, D0, dtimezero, temp, nmc, num_MC_update, move_timezero, k, temp_end=None):
        self.dv, self.dw, self.dwrad, self.D0, self.dtimezero = dv, dw, dwrad, D0, dtimezero
        self.temp = self.temp_start = temp
        self.temp_end = temp_end
        self.nmc = nmc
        self.num_MC_update = num_MC_update
        self.move_timezero = move_timezero
        if self.temp_end is None:
            self.temp_end = temp
            self.dtemp = 0.0
        else:
            nupdate = self.nmc / self.num_MC_update - 1 if self.num_MC_update else 0
            self.dtemp = (self.temp_end - self.temp_start) / float(nupdate) if nupdate > 0 else 0.0
        self.naccv = self.naccw = self.naccwrad = self.nacctimezero = 0
        self.naccv_update = self.naccw_update = self.naccwrad_update = self.nacctimezero_update = 0
        self.k = k

    # ---- lib/MCState.py:63-99 ------------------------------------------------------------
    def set_model(self, model, data, ncosF, ncosD, ncosDrad, pull):
        self.data = data
        # pull (kBT/angstrom) -> dF (kBT) between bins, lib/MCState.py:65-69.  The generator is then not
        # symmetrisable; the library runs it on the general squaring engine (pbc only, as the reference:
        # lib/utils.py:49-53 has no non-periodic pull variant).
        self.pull = None if pull is None else -pull * self.data.dz
        if self.pull is not None and not self.pbc:
            raise NotImplementedError("pull without pbc (lib/utils.py:49-53)")
        if self.pull is not None and self.do_radial:
            raise NotImplementedError("--pull together with the radial model (--rad) is not supported by this build")
        if self.do_radial:
            if model == "RadCosinusModel":
                self.model = RadCosinusModel(self.data, self.D0, ncosF, ncosD, 0, ncosDrad)
            elif model == "RadModel":
                self.model = RadModel(self.data, self.D0, ncosF, ncosD, 0)
            else:
                raise ValueError("model %s not found" % model)
            from scipy import special
            zeros = special.jn_zeros(0, self.lmax)                       # lib/twod.py:45-59
            r = np.arange(self.model.dim_rad, dtype=np.float64) + 0.5
            rmax = np.float64(self.model.dim_rad)
            self.model.bessel0_zeros = zeros
            self.model.bessels = np.array([2 * r * special.j0(r / rmax * z) / special.j1(z) ** 2 / rmax ** 2
                                           for z in zeros])
        else:
            if model in MODELS_1D:
                self.model = MODELS_1D[model](self.data, self.D0, ncosF, ncosD, 0)
            elif model == "Model":
                self.model = Model(self.data, self.D0)
            else:
                raise ValueError("model %s not found" % model)
        if self.pbc != self.model.pbc:
            raise AssertionError("pbc flag and counting method of the transition matrices disagree")

    # ---- lib/MCState.py:139-201 ------------------------------------------------------------
    def use_initfile(self, initfile, final=True):
        m = self.model
        if m.ncosF > 0:
            c = read_Fcoeffs(initfile, final=True)
            if len(c) > 0:
                print("USING initfile for v_coeff", initfile, len(c), "coeffs")
                k = min(len(c), m.ncosF)
                m.v_coeff[:k] = c[:k]
                m.update_v()
        else:
            F, D, edges = read_F_D_edges(initfile)
            if len(F) != len(m.v):
                raise AssertionError("initfile has %d bins, model has %d" % (len(F), len(m.v)))
            print("USING initfile for v", initfile, len(F), "values")
            m.v = F
        if m.ncosD > 0:
            c = read_Dcoeffs(initfile, final=True)
            if len(c) > 0:
                print("USING initfile for w_coeff", initfile, len(c), "coeffs")
                k = min(len(c), m.ncosD)
                m.w_coeff[:k] = c[:k]
                m.w_coeff[0] -= m.wunit
                m.update_w()
        else:
            F, D, edges = read_F_D_edges(initfile)
            if len(D) != len(m.w):
                raise AssertionError("initfile has %d D values, model has %d" % (len(D), len(m.w)))
            print("USING initfile for w", initfile, len(D), "values")
            m.w = np.log(D) - m.wunit
        if self.do_radial:                    # lib/MCState.py:170-197
            if getattr(m, "ncosDrad", 0) > 0:
                c = read_Dradcoeffs(initfile, final=True)
                if len(c) == 0:
                    c = read_Dcoeffs(initfile, final=True)     # the reference falls back to w_coeff
                if len(c) > 0:
                    print("USING initfile for wrad_coeff", initfile, len(c), "coeffs")
                    k = min(len(c), m.ncosDrad)
                    m.wrad_coeff[:k] = c[:k]
                    m.wrad_coeff[0] -= m.wradunit
                    m.update_wrad()
            else:
                Drad, redges = read_Drad(initfile)
                if len(Drad) != len(m.wrad):
                    raise AssertionError("initfile has %d Drad values, model has %d" % (len(Drad), len(m.wrad)))
                print("USING initfile for wrad", initfile, len(Drad), "values")
                m.wrad = np.log(Drad) - m.wradunit
        read_dv_dw(initfile, final=True)      # read but not applied, like lib/MCState.py:199-201

    # ---- device set-up -----------------------------------------------------------------------
    def _ensure_device(self):
        if self.chains is not None:
            return
        from .engine import ChainBatch, DeviceProblem, Engine
        m = self.model
        self.engine = Engine(self.device)
        if self.do_radial:
            from .engine import RadialChains
            ncos = getattr(m, "ncosDrad", 0)
            self.chains = RadialChains(self.engine, m.v, m.w, self.pbc, self.data.list_lt, self.data.list_trans,
                                       m.bessels, m.bessel0_zeros, self.nchains,
                                       wrad_basis=m.wrad_basis if ncos > 0 else None)
            wrad = np.tile(m.wrad, (self.nchains, 1))
            wc = np.tile(m.wrad_coeff, (self.nchains, 1)) if ncos > 0 else None
            if self.jitter > 0:
                if ncos > 0:
                    wc += self._jitter_rows(2, ncos)
                    wrad = wc @ m.wrad_basis.T
                else:
                    wrad += self._jitter_rows(2, wrad.shape[1])
            self.chains.set_state(wrad, wc, dwrad=self.dwrad, temp=self.temp)
            self._apply_resume()
            return
        svecs = _string_vecs(m.dim_w, self.pbc) if self.k > 0 else None
        self.string_vecs = svecs
        counts = np.asarray(self.data.list_trans)
        if not np.all(counts == np.round(counts)):
            raise ValueError("transition counts must be integers")
        self.problem = DeviceProblem(self.engine, counts.astype(np.int64), self.data.list_lt, self.pbc,
                                     v_basis=m.v_basis if m.ncosF > 0 else None,
                                     w_basis=m.w_basis if m.ncosD > 0 else None,
                                     spring_k=self.k if self.k > 0 else -1.0, string_vecs=svecs, pull=self.pull)
        self.chains = ChainBatch(self.problem, self.nchains)
        v = np.tile(m.v, (self.nchains, 1))
        w = np.tile(m.w, (self.nchains, 1))
        vc = np.tile(m.v_coeff, (self.nchains, 1)) if m.ncosF > 0 else None
        wc = np.tile(m.w_coeff, (self.nchains, 1)) if m.ncosD > 0 else None
        if self.jitter > 0:     # optional over-dispersed starts; global chain 0 keeps the exact start
            if m.ncosF > 0:
                vc[:, 1:] += self._jitter_rows(0, m.ncosF)[:, 1:]
                v = vc @ m.v_basis.T
            else:
                v += self._jitter_rows(0, v.shape[1])
            if m.ncosD > 0:
                wc += self._jitter_rows(1, m.ncosD)
                w = wc @ m.w_basis.T
            else:
                w += self._jitter_rows(1, w.shape[1])
        self.chains.set_state(v, w, vc, wc, dv=self.dv, dw=self.dw, dtimezero=self.dtimezero, temp=self.temp,
                              timezero=m.timezero)
        self._apply_resume()

    def _jitter_rows(self, stream, width):
        """Gaussian start offsets of the local chains, [nchains][width].  Row c depends only on (seed, stream,
        GLOBAL chain index): every rank draws different rows and a run does not depend on the number of GPUs.
        Global chain 0 gets zeros (it keeps the exact start of the reference's single chain)."""
        out = np.zeros((self.nchains, width))
        for c in range(self.nchains):
            g = self.chain_lo + c
            if g == 0:
                continue
            rng = np.random.default_rng([self.seed & 0xFFFFFFFFFFFFFFFF, 0x5DEECE66D, int(stream), int(g)])
            out[c] = self.jitter * rng.standard_normal(width)
        return out

    # ---- exact resume: <outfile>.chains.npz of an earlier run ----------------------
    def use_chainfile(self, filename):
        """Continue every chain from the state an earlier run left in `<outfile>.chains.npz`: profiles,
        coefficients, step widths, temperature, log-likelihood, acceptance counters and the step counter
        (= Philox counter; the streams are keyed by (seed, global chain index, step), so with the same --seed the
        continued run draws exactly the numbers the uninterrupted run would have drawn)."""
        z = np.load(filename)
        if int(z["scal"].shape[0]) != self.nchains_total:
            raise ValueError("%s holds %d chains, this run has %d" % (filename, z["scal"].shape[0], self.nchains_total))
        if bool(int(z["radial"])) != bool(self.do_radial):
            raise ValueError("%s is a %s run" % (filename, "radial" if int(z["radial"]) else "1-D"))
        self._resume = {k: z[k] for k in z.files}
        self.seed = int(z["seed"])
        self.steps_done = int(z["steps_done"])

    def _apply_resume(self):
        z = getattr(self, "_resume", None)
        if z is None:
            return
        import torch
        lo, hi = self.chain_lo, self.chain_hi
        ch, eng = self.chains, self.engine
        f64, i64 = torch.float64, torch.int64

        def put(dst, a, dtype):
            a = np.ascontiguousarray(a[lo:hi])
            if a.shape != tuple(dst.shape):
                raise ValueError("chain file does not match this model: %s vs %s" % (a.shape, tuple(dst.shape)))
            dst.copy_(eng.to_device(a, dtype))

        with torch.cuda.device(eng.device):
            put(ch.scal, z["scal"], f64)
            put(ch.counters, z["counters"], i64)
            if self.do_radial:
                put(ch.wrad, z["wrad"], f64)
                if ch.ncos > 0:
                    put(ch.wcoef, z["wrad_coeff"], f64)
                    put(ch.nacc_coeff, z["naccwrad_coeff"], i64)
            else:
                put(ch.v, z["v"], f64)
                put(ch.w, z["w"], f64)
                if self.model.ncosF > 0:
                    put(ch.vcoef, z["v_coeff"], f64)
                    put(ch.naccv_coeff, z["naccv_coeff"], i64)
                if self.model.ncosD > 0:
                    put(ch.wcoef, z["w_coeff"], f64)
                    put(ch.naccw_coeff, z["naccw_coeff"], i64)

    # ---- lib/MCState.py:101-137 ----------------------------------------------------------------
    def init_log_like(self):
        self._ensure_device()
        keep = self.chains.scal[:, 0].clone() if getattr(self, "_resume", None) is not None else None
        ll = self.chains.init_log_like()          # raises ValueError("Initial likelihood diverges") on NaN
        if keep is not None:                      # resumed run: the running value of the chain, not a recomputation
            self.chains.scal[:, 0] = keep
            ll = to_host(keep)
        self.chain_log_like = ll
        self.log_like = float(ll[0]) if len(ll) else float("nan")   # a rank may own no chain (world > nchains)
        print("initial log-likelihood:", self.log_like)
        m = self.model
        if m.ncosF > 0:
            self.naccv_coeff = np.zeros(m.ncosF, int)
        if m.ncosD > 0:
            self.naccw_coeff = np.zeros(m.ncosD, int)
        if self.do_radial and getattr(m, "ncosDrad", 0) > 0:
            self.naccwrad_coeff = np.zeros(m.ncosDrad, int)
        return ll

    # ---- the hot loop: lib/mc.py:27-51 on the GPU ------------------------------------------------
    def mc_params(self):
        mp = K.McParams()
        mp.num_mc_update = float(self.num_MC_update)
        mp.dtemp = float(self.dtemp)
        mp.min_lt = float(self.data.min_lt)
        mp.move_timezero = int(bool(self.move_timezero))
        mp.sample_every = 100
        mp.refresh_every = 512
        mp.use_tape = 0
        mp.method = self.method
        mp.taylor_m = 0
        mp.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        mp.chain_offset = self.chain_lo
        mp.tape_stride = 0
        return mp

    def advance(self, nsteps, want_samples=True):
        """Run nsteps MC steps on every chain; returns the host sample rows
        [nchains][k][sample_dim] taken at steps that are multiples of 100."""
        self._ensure_device()
        first = self.steps_done
        nsamp = (first + nsteps) // 100 - first // 100 if want_samples else 0
        if self.do_radial:
            _, _, smp = self.chains.run(nsteps, num_mc_update=float(self.num_MC_update), dtemp=float(self.dtemp),
                                        seed=self.seed & 0xFFFFFFFFFFFFFFFF, chain_offset=self.chain_lo,
                                        nsamples_cap=nsamp)
            self.steps_done += nsteps
            self.sync_from_device()
            return None if smp is None else to_host(smp)
        _, _, smp = self.chains.run(nsteps, self.mc_params(), nsamples_cap=nsamp)
        self.steps_done += nsteps
        self.sync_from_device()
        return None if smp is None else to_host(smp)

    def sync_from_device(self):
        self.set_host_state(self.chains.host_state())

    def set_host_state(self, st):
        """Per-chain arrays + the reference's scalar attributes (from row 0) out of a host_state() dictionary --
        this rank's chains, or the rows of ALL chains gathered on rank 0 (infty.run_lag_scan)."""
        self.chain_scal, self.chain_counters = st["scal"], st["counters"]
        if self.do_radial:
            m = self.model
            sc, ct = st["scal"], st["counters"]
            self.chain_naccwrad_coeff = st["nacc_coeff"]
            self.chain_wrad, self.chain_wrad_coeff = st["wrad"], st["wcoef"]
            self.chain_log_like = sc[:, 0].copy()
            self.chain_dwrad, self.chain_temp = sc[:, 1].copy(), sc[:, 2].copy()
            self.chain_naccwrad, self.chain_status = ct[:, 0].copy(), ct[:, 3].copy()
            if len(sc) == 0:
                return
            m.wrad = st["wrad"][0].copy()
            if getattr(m, "ncosDrad", 0) > 0:
                m.wrad_coeff = st["wcoef"][0].copy()
                self.naccwrad_coeff = st["nacc_coeff"][0].copy()
            self.log_like = float(sc[0, 0])
            self.dwrad, self.temp = float(sc[0, 1]), float(sc[0, 2])
            self.naccwrad = int(ct[0, 0])
            return
        self.chain_v, self.chain_w = st["v"], st["w"]
        self.chain_v_coeff, self.chain_w_coeff = st["vcoef"], st["wcoef"]
        sc, ct = st["scal"], st["counters"]
        self.chain_log_like = sc[:, K.S_LL].copy()
        self.chain_timezero = sc[:, K.S_T0].copy()
        self.chain_dv, self.chain_dw = sc[:, K.S_DV].copy(), sc[:, K.S_DW].copy()
        self.chain_temp = sc[:, K.S_TEMP].copy()
        self.chain_naccv, self.chain_naccw = ct[:, K.C_NACCV].copy(), ct[:, K.C_NACCW].copy()
        self.chain_nacctimezero = ct[:, K.C_NACCT0].copy()
        self.chain_status = ct[:, K.C_STATUS].copy()
        self.chain_naccv_coeff, self.chain_naccw_coeff = st["naccv_coeff"], st["naccw_coeff"]
        if len(sc) == 0:
            return
        # chain 0 -> the reference's scalar attributes
        m = self.model
        m.v, m.w = st["v"][0].copy(), st["w"][0].copy()
        if m.ncosF > 0:
            m.v_coeff = st["vcoef"][0].copy()
            self.naccv_coeff = st["naccv_coeff"][0].copy()
        if m.ncosD > 0:
            m.w_coeff = st["wcoef"][0].copy()
            self.naccw_coeff = st["naccw_coeff"][0].copy()
        m.timezero = float(sc[0, K.S_T0])
        m.list_lt = np.asarray(self.data.list_lt) + m.timezero
        self.log_like = float(sc[0, K.S_LL])
        self.dv, self.dw, self.temp = float(sc[0, K.S_DV]), float(sc[0, K.S_DW]), float(sc[0, K.S_TEMP])
        self.naccv, self.naccw = int(ct[0, K.C_NACCV]), int(ct[0, K.C_NACCW])
        self.nacctimezero = int(ct[0, K.C_NACCT0])

    # ---- printing (same text as lib/MCState.py:396-491) ----------------------------------------
    def print_MC_params(self, f=None, final=False):
        f = f or sys.stdout
        print("----- final Settings MC -----" if final else "----- Settings MC -----", file=f)
        print("dv(MC-potential)=", self.dv, file=f)
        print("dw(MC-logD)=", self.dw, file=f)
        print("dwrad(MC-logDrad)=", self.dwrad, file=f)
        print("temp=", self.temp, file=f)
        print("n(MC)=", self.nmc, file=f)
        print("n(update)=", self.num_MC_update, file=f)
        print("k=", self.k, file=f)
        print("-" * 20, file=f)

    def print_intermediate(self, imc, printfreq=100):
        step = imc + 1
        print(imc, self.log_like, float(self.naccv) / step, float(self.naccw) / step, float(self.naccwrad) / step)

    def print_statistics(self, f=None):
        f = f or sys.stdout
        print("===== Statistics =====", file=f)
        print("nmc         ", self.nmc, file=f)
        print("naccv       ", self.naccv, file=f)
        print("naccw       ", self.naccw, file=f)
        print("naccwrad    ", self.naccwrad, file=f)
        print("nacctimezero", self.nacctimezero, file=f)
        for name, val in (("accv ratio       ", self.naccv), ("accw ratio       ", self.naccw),
                          ("accwrad ratio    ", self.naccwrad), ("acctimezero ratio", self.nacctimezero)):
            print(name, "%5.1f" % (float(val) / self.nmc * 100), "%", file=f)
        print("=" * 10, file=f)
        m = self.model
        for ncos, title, arr in ((m.ncosF, "naccv_coeff", getattr(self, "naccv_coeff", None)),
                                 (m.ncosD, "naccw_coeff", getattr(self, "naccw_coeff", None))):
            if ncos > 0:
                tot = max(1, np.sum(arr))
                print(title, file=f)
                for i, val in enumerate(arr):
                    print("%8d %8d %5.1f %s %5.1f %s" % (i, val, float(val) / tot * 100, "%",
                                                        float(val) / self.nmc * 100, "%"), file=f)
        if self.do_radial and getattr(m, "ncosDrad", 0) > 0:
            tot = max(1, np.sum(self.naccwrad_coeff))
            print("naccwrad_coeff", file=f)
            for i, val in enumerate(self.naccwrad_coeff):
                print("%8d %8d %5.1f %s %5.1f %s" % (i, val, float(val) / tot * 100, "%",
                                                    float(val) / self.nmc * 100, "%"), file=f)
        if self.nchains_total > 1:
            g = getattr(self, "global_summary", None)
            if g is not None:        # after the gather: every chain of every rank
                acc, ll = g["acc"], g["log_like"]
            else:
                acc = ((self.chain_naccwrad if self.do_radial else self.chain_naccv + self.chain_naccw)
                       / float(max(1, self.nmc)))
                ll = self.chain_log_like
            print("===== Statistics over %d chains =====" % len(ll), file=f)
            print("acceptance mean/min/max %6.3f %6.3f %6.3f" % (acc.mean(), acc.min(), acc.max()), file=f)
            print("log-like   mean/min/max", ll.mean(), ll.min(), ll.max(), file=f)

    def set_global_summary(self, log_like, acc):
        """Per-chain log-likelihood and acceptance ratio of ALL chains (after the final all_gather)."""
        self.global_summary = dict(log_like=np.asarray(log_like), acc=np.asarray(acc))

    def print_coeffs_laststate(self, f, final=False):
        from .log import print_coeffs
        m = self.model
        wrad_coeff = m.wrad_coeff if (self.do_radial and getattr(m, "ncosDrad", 0) > 0) else None
        print_coeffs(f, m, m.v_coeff if m.ncosF > 0 else None, m.w_coeff if m.ncosD > 0 else None, wrad_coeff,
                     m.timezero if self.move_timezero else None, final=final)

    def print_laststate(self, f, final=False):
        from .log import print_profiles
        self.print_MC_params(f, final=final)
        self.print_coeffs_laststate(f, final=final)
        print_profiles(f, self.model, self.model.v, self.model.w, wrad=self.model.wrad if self.do_radial else None,
                       final=final)

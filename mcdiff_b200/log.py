"""Sampling log and result writers.

`Logger` keeps the rows the reference's Logger stores every 100 MC steps (lib/log.py:6-58):
log_like, timezero, dv, dw, dwrad, dtimezero and the sampled parameters (v or v_coeff, w or
w_coeff, wrad or wrad_coeff).  With many chains the rows of all chains are pooled (chain-major),
so every statistic taken over axis 0 -- as lib/outreading.py:281-301 does -- is the pooled
posterior estimate.  The text writers reproduce the reference's formats exactly
(lib/log.py:288-376); they are the files `outreading`, `plotresults`, `mcdiff-infty` and
`--initf` parse.
"""
from __future__ import annotations

import pickle
import sys

import numpy as np


class Logger:
    freq = 100

    def __init__(self, MC=None, nchains=1):
        self.nchains = int(nchains)
        if MC is None:
            return
        self.nmc = MC.nmc
        self.freq = 100
        self.nf_chain = MC.nmc // self.freq + 1       # rows per chain, initial state included
        self.nf = self.nf_chain * self.nchains
        m = MC.model
        nf = self.nf
        self.log_like = np.zeros(nf)
        self.timezero = np.zeros(nf)
        self.dv = np.zeros(nf)
        self.dw = np.zeros(nf)
        self.dwrad = np.zeros(nf)
        self.dtimezero = np.zeros(nf)
        self.temp = np.zeros(nf)
        if m.ncosF <= 0:
            self.v = np.zeros((nf, m.dim_v))
        else:
            self.v_coeff = np.zeros((nf, m.ncosF))
        if m.ncosD <= 0:
            self.w = np.zeros((nf, m.dim_w))
        else:
            self.w_coeff = np.zeros((nf, m.ncosD))
        if getattr(MC, "do_radial", False):
            if m.ncosDrad <= 0:
                self.wrad = np.zeros((nf, m.dim_wrad))
            else:
                self.wrad_coeff = np.zeros((nf, m.ncosDrad))

    # ---- filling from the device sample buffers -----------------------------------------
    def fill_rows(self, chain, first_row, rows, model):
        """rows: [k][sample_dim] = {log_like, timezero, dv, dw, dtimezero, temp, pv..., pw...}
        written to rows first_row.. of `chain`."""
        k = rows.shape[0]
        if k == 0:
            return
        a = chain * self.nf_chain + first_row
        b = a + k
        self.log_like[a:b] = rows[:, 0]
        self.timezero[a:b] = rows[:, 1]
        self.dv[a:b] = rows[:, 2]
        self.dw[a:b] = rows[:, 3]
        self.dtimezero[a:b] = rows[:, 4]
        self.temp[a:b] = rows[:, 5]
        nv = model.ncosF if model.ncosF > 0 else model.dim_v
        pv, pw = rows[:, 6:6 + nv], rows[:, 6 + nv:]
        if model.ncosF > 0:
            self.v_coeff[a:b] = pv
        else:
            self.v[a:b] = pv
        if model.ncosD > 0:
            self.w_coeff[a:b] = pw
        else:
            self.w[a:b] = pw

    def fill_rows_all(self, first_row, samples, model):
        """samples: [nchains][k][sample_dim] of ALL chains of this Logger at once (one strided copy per field instead
        of a Python call per chain: 1024 chains cost 3 ms per chunk the slow way)."""
        k = samples.shape[1]
        if k == 0:
            return
        a, b = first_row, first_row + k

        def view(arr):
            return arr.reshape((self.nchains, self.nf_chain) + arr.shape[1:])
        for j, name in enumerate(("log_like", "timezero", "dv", "dw", "dtimezero", "temp")):
            view(getattr(self, name))[:, a:b] = samples[:, :, j]
        nv = model.ncosF if model.ncosF > 0 else model.dim_v
        view(self.v_coeff if model.ncosF > 0 else self.v)[:, a:b] = samples[:, :, 6:6 + nv]
        view(self.w_coeff if model.ncosD > 0 else self.w)[:, a:b] = samples[:, :, 6 + nv:]

    def fill_rows_radial(self, chain, first_row, rows, MC):
        """rows: [k][3 + nw] = {log_like, dwrad, temp, wrad | wrad_coeff}; v, w are frozen in radial runs."""
        k = rows.shape[0]
        if k == 0:
            return
        m = MC.model
        a = chain * self.nf_chain + first_row
        b = a + k
        self.log_like[a:b] = rows[:, 0]
        self.dwrad[a:b] = rows[:, 1]
        self.temp[a:b] = rows[:, 2]
        self.dv[a:b], self.dw[a:b], self.dtimezero[a:b] = MC.dv, MC.dw, MC.dtimezero
        if m.ncosF > 0:
            self.v_coeff[a:b] = m.v_coeff
        else:
            self.v[a:b] = m.v
        if m.ncosD > 0:
            self.w_coeff[a:b] = m.w_coeff
        else:
            self.w[a:b] = m.w
        if getattr(m, "ncosDrad", 0) > 0:
            self.wrad_coeff[a:b] = rows[:, 3:]
        else:
            self.wrad[a:b] = rows[:, 3:]

    def chain_rows(self, chain):
        return slice(chain * self.nf_chain, (chain + 1) * self.nf_chain)

    # ---- pooling over ranks: every stored quantity of the local chains as one [nchains][nf_chain * K] block ----
    _SCALARS = ("log_like", "timezero", "dv", "dw", "dwrad", "dtimezero", "temp")
    _VECTORS = ("v", "v_coeff", "w", "w_coeff", "wrad", "wrad_coeff")

    def pack_rows(self):
        """All rows of the local chains, chain-major: [nchains][nf_chain * K].  Column order: the scalars of
        `_SCALARS`, then every sampled vector this Logger holds (order of `_VECTORS`)."""
        cols = [getattr(self, k).reshape(self.nchains, self.nf_chain, 1) for k in self._SCALARS]
        for k in self._VECTORS:
            if hasattr(self, k):
                a = getattr(self, k)
                cols.append(a.reshape(self.nchains, self.nf_chain, a.shape[1]))
        width = sum(c.shape[2] for c in cols)
        return np.ascontiguousarray(np.concatenate(cols, axis=2).reshape(self.nchains, self.nf_chain * width))

    def unpack_rows(self, rows):
        """Inverse of pack_rows for a Logger sized for rows.shape[0] chains (global chain order)."""
        rows = np.asarray(rows, dtype=np.float64)
        rows = rows.reshape(self.nchains, self.nf_chain, rows.shape[1] // self.nf_chain)
        c = 0
        for k in self._SCALARS:
            getattr(self, k)[:] = rows[:, :, c].reshape(-1)
            c += 1
        for k in self._VECTORS:
            if hasattr(self, k):
                a = getattr(self, k)
                a[:] = rows[:, :, c:c + a.shape[1]].reshape(a.shape)
                c += a.shape[1]
        if c != rows.shape[2]:
            raise ValueError("row width %d does not match this Logger (%d)" % (rows.shape[2], c))

    # ---- pickling (lib/log.py:68-71, 279-284) ---------------------------------------------
    def dump(self, filename):
        with open(filename, "wb") as f:
            pickle.dump(self, f)

    # ---- statistics -------------------------------------------------------------------------
    def _profiles(self, coeff, basis):
        return np.asarray(coeff) @ np.asarray(basis).T        # rows: sum_k coeff_k basis[:, k]

    # Posterior statistics of many chains: 8192 chains x 11 rows x 100 bins are 9 M profile values per quantity, and
    # the matrix product, the exponentials and the four reductions cost rank 0 ~0.2 s on the host -- as long as the
    # whole sampling.  With a CUDA device at hand they are formed there (same formulas: population std, mean of e^w).
    _DEVICE_ROWS = 16384

    def _device_stats(self, coeff, basis, want_log, want_exp):
        """(mean, std) of the profile rows and of their exponentials on the GPU, or None when no device / few rows."""
        coeff = np.asarray(coeff)
        if coeff.shape[0] < self._DEVICE_ROWS:
            return None
        try:
            import torch
            if not torch.cuda.is_available():
                return None
            dev = torch.device("cuda", torch.cuda.current_device())
            a = torch.from_numpy(np.ascontiguousarray(coeff, dtype=np.float64)).to(dev) @ \
                torch.from_numpy(np.ascontiguousarray(basis, dtype=np.float64)).to(dev).T
            out = [None, None, None, None]
            if want_log:
                out[0], out[1] = a.mean(0).cpu().numpy(), a.std(0, unbiased=False).cpu().numpy()
            if want_exp:
                ea = torch.exp(a)
                out[2], out[3] = ea.mean(0).cpu().numpy(), ea.std(0, unbiased=False).cpu().numpy()
            return tuple(out)
        except Exception:
            return None

    def average_profile_v_from_coeff(self):
        d = self._device_stats(self.v_coeff, self.model.v_basis, True, False)
        if d is not None:
            return d[0], d[1]
        a = self._profiles(self.v_coeff, self.model.v_basis)
        return np.mean(a, 0), np.std(a, 0)

    def average_profile_w_from_coeff(self, want_log=True):
        """(mean w, std w, mean e^w, std e^w) over all stored rows; want_log=False skips the first two (None, None):
        the result files only need the statistics of D = e^w, and every reduction is a pass over 9 M values at 8192
        chains."""
        d = self._device_stats(self.w_coeff, self.model.w_basis, want_log, True)
        if d is not None:
            return d
        a = self._profiles(self.w_coeff, self.model.w_basis)
        ea = np.exp(a)
        lm, ls = (np.mean(a, 0), np.std(a, 0)) if want_log else (None, None)
        return lm, ls, np.mean(ea, 0), np.std(ea, 0)

    def average_profile_wrad_from_coeff(self, want_log=True):
        d = self._device_stats(self.wrad_coeff, self.model.wrad_basis, want_log, True)
        if d is not None:
            return d
        a = self._profiles(self.wrad_coeff, self.model.wrad_basis)
        ea = np.exp(a)
        lm, ls = (np.mean(a, 0), np.std(a, 0)) if want_log else (None, None)
        return lm, ls, np.mean(ea, 0), np.std(ea, 0)

    def print_MC_params(self, f=None, final=False):
        f = f or sys.stdout
        print("----- final Settings MC -----" if final else "----- Settings MC -----", file=f)
        print("dv(MC-potential)=", self.dv[-1], file=f)
        print("dw(MC-logD)=", self.dw[-1], file=f)
        print("dwrad(MC-logDrad)=", self.dwrad[-1], file=f)
        print("n(MC)=", self.nmc, file=f)
        print("-" * 20, file=f)

    def print_average(self, model, f=None):
        """Posterior mean +- std in the `.ave.dat` layout (lib/log.py:118-128)."""
        from .outreading import read_F_D_edges_logger, read_Drad_logger, read_coeff_logger
        f = f or sys.stdout
        F, D, edges, Fst, Dst = read_F_D_edges_logger(self)
        Drad, redges, Dradst = read_Drad_logger(self)
        v_coeff, w_coeff, wrad_coeff, _, _, _, timezero, _ = read_coeff_logger(self)
        self.print_MC_params(f=f, final=True)
        print_coeffs(f, model, v_coeff, w_coeff, wrad_coeff, timezero, final=True)
        print_profiles(f, model, F, D, Drad, final=True, error=[Fst, Dst, Dradst], unit="notinternal")

    def statistics(self, MC, st=0, f=None):
        """mean/std table of every sampled quantity (lib/log.py:130-180)."""
        f = f or sys.stdout
        s = st // self.freq
        m = MC.model

        def table(vec):
            print("VEC", vec.shape, file=f)
            for i in range(vec.shape[1]):
                print(i, np.mean(vec[s:, i]), np.std(vec[s:, i]), file=f)

        print("===== stat v =====", file=f)
        if m.ncosF <= 0:
            table(self.v)
        else:
            table(self._profiles(self.v_coeff, m.v_basis))
            print("===== stat v_coeff =====", file=f)
            table(self.v_coeff)
        print("===== stat w =====", file=f)
        if m.ncosD <= 0:
            table(self.w)
        else:
            table(self._profiles(self.w_coeff, m.w_basis))
            print("===== stat w_coeff =====", file=f)
            table(self.w_coeff)


def load_logger(filename):
    with open(filename, "rb") as f:
        return pickle.load(f)


def print_coeffs(f, model, v_coeff=None, w_coeff=None, wrad_coeff=None, timezero=None, final=False):
    """Coefficient tables; the flat w coefficient is printed shifted by +wunit (lib/log.py:288-312)."""
    tag = "final " if final else ""
    if model.ncosF > 0:
        print("===== %sv_coeff =====" % tag, file=f)
        for i, val in enumerate(v_coeff):
            print("%8d %16.8e" % (i, val), file=f)
    if model.ncosD > 0:
        print("===== %sw_coeff =====" % tag, file=f)
        print("%8d %16.8e" % (0, w_coeff[0] + model.wunit), file=f)
        for i, val in enumerate(w_coeff[1:]):
            print("%8d %16.8e" % (i + 1, val), file=f)
    if timezero is not None:
        print("===== %stimezero =====" % tag, file=f)
        print("%16.8e" % timezero, file=f)
    if wrad_coeff is not None and getattr(model, "ncosDrad", 0) > 0:
        print("===== %swrad_coeff =====" % tag, file=f)
        print("%8d %16.8e" % (0, wrad_coeff[0] + model.wradunit), file=f)
        for i, val in enumerate(wrad_coeff[1:]):
            print("%8d %16.8e" % (i + 1, val), file=f)
    print("=" * 10, file=f)


def print_profiles(f, model, v, w, wrad=None, final=False, error=None, unit="internal"):
    """F / D table (lib/log.py:316-376).  unit="internal": w, wrad are logs in internal units and
    converted with wunit / wradunit; otherwise they already are D values in A^2/ps."""
    print("===== final F D =====" if final else "===== F D =====", file=f)
    F = v
    edges = model.edges
    if unit == "internal":
        D = np.exp(w + model.wunit)
        Drad = np.exp(wrad + model.wradunit) if wrad is not None else None
    else:
        D, Drad = w, wrad
    Fst = Dst = Dradst = None
    if error is not None:
        Fst, Dst, Dradst = error
        if unit == "internal":
            Dst = np.exp(error[1] + model.wunit)
            if wrad is not None:
                Dradst = np.exp(error[2] + model.wradunit)
    f.write("%8s %8s %8s  %13s %s" % ("index", "bin-str", "bin-end", "potential",
                                     "diffusion-coefficient(shifted-by-half-bin)\n"))
    maxi = model.dim_v if model.pbc else model.dim_v - 1
    for i in range(maxi):
        if error is None:
            f.write("%8d %8.3f %8.3f  %16.8e %16.8e\n" % (i, edges[i], edges[i + 1], F[i], D[i]))
        else:
            f.write("%8d %8.3f %8.3f  %16.8e %16.8e  %16.8e %16.8e\n" % (i, edges[i], edges[i + 1], F[i], D[i],
                                                                       Fst[i], Dst[i]))
    if not model.pbc:      # last bin has an F but no D (w lives between bins)
        i = model.dim_v - 1
        if error is None:
            f.write("%8d %8.3f %8.3f  %16.8e\n" % (i, edges[-2], edges[-1], F[-1]))
        else:
            f.write("%8d %8.3f %8.3f  %16.8e  %16.8e\n" % (i, edges[-2], edges[-1], F[-1], Fst[-1]))
    f.write("#Done\n")
    if wrad is not None:
        print("===== final Drad =====" if final else "===== Drad =====", file=f)
        f.write("%8s %8s %8s  %s" % ("index", "bin-str", "bin-end", "diffusion-coefficient-at[i]\n"))
        for i in range(model.dim_wrad):
            if error is None:
                f.write("%8d %8.3f %8.3f  %16.8e\n" % (i, edges[i], edges[i + 1], Drad[i]))
            else:
                f.write("%8d %8.3f %8.3f  %16.8e  %16.8e\n" % (i, edges[i], edges[i + 1], Drad[i], Dradst[i]))
        f.write("#Done\n")


def write_average_from_pic(picfilename, datfilename):
    """`.pic` -> `.pic.ave.dat` (lib/log.py:380-387)."""
    logger = load_logger(picfilename)
    with open(datfilename, "w") as f:
        logger.print_average(logger.model, f=f)

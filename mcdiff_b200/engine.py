"""Thin host layer over the C ABI: device buffers are torch tensors, compute is CUDA.

`Engine` owns one library handle per GPU, `DeviceProblem` the device copy of what the
reference keeps in `Transitions` + `Model` (count matrices, lag times, basis functions), and
`ChainBatch` the structure-of-arrays state of many independent Monte-Carlo chains.

There is no CPU path: every method ends in a call into libmcdiff_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi as K

SAMPLE_HEAD = 6  # log_like, timezero, dv, dw, dtimezero, temp

# bytes this layer moved between host and device (bench.py's e2e leg reports them per step)
TRAFFIC = {"h2d": 0, "d2h": 0}


def to_host(t):
    """Device tensor -> numpy (synchronising copy), counted in TRAFFIC."""
    TRAFFIC["d2h"] += t.numel() * t.element_size()
    return t.cpu().numpy()


def _dev(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, device: int | None = None):
        if not torch.cuda.is_available():
            raise K.McdError("mcdiff_b200 needs a CUDA device (no CPU fallback)")
        self.lib = K.load()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        K.check(self.lib.mcd_create(self.device_index, C.byref(h)), "mcd_create")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mcd_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stream_ptr(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def to_device(self, a, dtype):
        """Host array -> device tensor through pinned memory (async on the current stream)."""
        if isinstance(a, torch.Tensor):
            if not a.is_cuda:
                TRAFFIC["h2d"] += a.numel() * a.element_size()
            return a.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()
        a = np.ascontiguousarray(a)
        if not a.flags.writeable:
            a = a.copy()
        t = torch.from_numpy(a)
        if t.dtype != dtype:
            t = t.to(dtype)
        TRAFFIC["h2d"] += t.numel() * t.element_size()
        if t.numel() * t.element_size() < (1 << 20):
            return t.to(self.device)          # small: page-locking a fresh buffer costs more (~2 ms) than the staged copy
        return t.pin_memory().to(self.device, non_blocking=True)

    def fp64_peak_tflops(self) -> float:
        out = C.c_double()
        K.check(self.lib.mcd_fp64_peak(self.handle, C.byref(out)), "mcd_fp64_peak")
        return out.value

    def launch_count(self) -> int:
        return int(self.lib.mcd_launch_count(self.handle))


class DeviceProblem:
    """Device-resident problem description (counts, lag times, bases)."""

    def __init__(self, eng: Engine, counts, lagtimes, pbc, v_basis=None, w_basis=None, spring_k=-1.0,
                 string_vecs=None, clip=1e-10, pull=None):
        """pull: dF between neighbouring bins (= -pull_force * dz, lib/MCState.py:65-67) or None."""
        self.eng = eng
        self.pull = None if pull is None else float(pull)
        if self.pull is not None and not pbc:
            raise NotImplementedError("pull without pbc (lib/utils.py:49-53)")
        counts = np.asarray(counts)
        if counts.ndim != 3 or counts.shape[1] != counts.shape[2]:
            raise ValueError("counts must be [nlag][n][n]")
        if np.any(counts < 0) or np.any(counts > np.iinfo(np.int32).max):
            raise ValueError("counts out of int32 range")
        self.n = int(counts.shape[1])
        self.nlag = int(counts.shape[0])
        self.pbc = bool(pbc)
        self.dim_w = self.n if self.pbc else self.n - 1
        self.ncosF = 0 if v_basis is None else int(np.asarray(v_basis).shape[1])
        self.ncosD = 0 if w_basis is None else int(np.asarray(w_basis).shape[1])
        self.spring_k = float(spring_k)
        with torch.cuda.device(eng.device):
            self.d_counts = eng.to_device(counts.astype(np.int32), torch.int32)
            self.d_lag = eng.to_device(np.asarray(lagtimes, dtype=np.float64), torch.float64)
            self.d_vb = None if v_basis is None else eng.to_device(np.asarray(v_basis, dtype=np.float64), torch.float64)
            self.d_wb = None if w_basis is None else eng.to_device(np.asarray(w_basis, dtype=np.float64), torch.float64)
            self.d_sv = None
            if self.spring_k > 0:
                if string_vecs is None:
                    raise ValueError("spring_k > 0 needs string_vecs")
                self.d_sv = eng.to_device(np.asarray(string_vecs, dtype=np.float64), torch.float64)
            d = K.ProblemDesc()
            d.n, d.pbc, d.nlag, d.ncosF, d.ncosD = self.n, int(self.pbc), self.nlag, self.ncosF, self.ncosD
            d.lagtimes = self.d_lag.data_ptr()
            d.counts = self.d_counts.data_ptr()
            d.v_basis = None if self.d_vb is None else self.d_vb.data_ptr()
            d.w_basis = None if self.d_wb is None else self.d_wb.data_ptr()
            d.string_vecs = None if self.d_sv is None else self.d_sv.data_ptr()
            d.spring_k, d.clip = self.spring_k, float(clip)
            d.use_pull, d.pull = (0, 0.0) if self.pull is None else (1, self.pull)
            p = C.c_void_p()
            K.check(eng.lib.mcd_problem_create(eng.handle, C.byref(d), eng.stream_ptr(), C.byref(p)),
                    "mcd_problem_create")
            self.ptr = p
            np_, be, sd, ins, sq = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
            K.check(eng.lib.mcd_problem_info(p, C.byref(np_), C.byref(be), C.byref(sd), C.byref(ins), C.byref(sq)))
            self.np, self.basis_elems, self.sample_dim, self.in_smem = np_.value, be.value, sd.value, bool(ins.value)
            self.squaring_ok = bool(sq.value)
        self.lagtimes = np.asarray(lagtimes, dtype=np.float64).copy()
        self.min_lt = float(np.min(self.lagtimes))

    def close(self):
        if getattr(self, "ptr", None):
            self.eng.lib.mcd_problem_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- likelihood seam: mcdiff.utils.log_like_lag (lib/utils.py:355-368) ----
    def loglike_device(self, d_v, d_w, d_lag=None, want_P=False, method=K.METHOD_AUTO):
        """Device tensors in, device tensors out (asynchronous)."""
        eng = self.eng
        B = d_v.shape[0]
        out = torch.empty(B, dtype=torch.float64, device=eng.device)
        status = torch.empty(B, dtype=torch.int32, device=eng.device)
        P = torch.empty((B, self.nlag, self.n, self.n), dtype=torch.float64, device=eng.device) if want_P else None
        K.check(eng.lib.mcd_loglike_batch(eng.handle, self.ptr, int(method), B, _dev(d_v), _dev(d_w), _dev(d_lag), _dev(out),
                                          _dev(P), _dev(status), eng.stream_ptr()), "mcd_loglike_batch")
        return out, P, status

    def loglike(self, v, w, lagtimes=None, want_P=False, method=K.METHOD_AUTO):
        """Host arrays in, host arrays out: log L (and optionally the propagators) of B profiles."""
        eng = self.eng
        v = np.atleast_2d(np.asarray(v, dtype=np.float64))
        w = np.atleast_2d(np.asarray(w, dtype=np.float64))
        if v.shape[1] != self.n or w.shape[1] != self.dim_w or v.shape[0] != w.shape[0]:
            raise ValueError("profile shapes do not match the problem")
        with torch.cuda.device(eng.device):
            d_v = eng.to_device(v, torch.float64)
            d_w = eng.to_device(w, torch.float64)
            d_l = None
            if lagtimes is not None:
                lt = np.asarray(lagtimes, dtype=np.float64)
                if lt.shape != (v.shape[0], self.nlag):
                    raise ValueError("lagtimes must be [B][nlag]")
                d_l = eng.to_device(lt, torch.float64)
            out, P, status = self.loglike_device(d_v, d_w, d_l, want_P, method)
            ll = to_host(out)
            st = to_host(status)
            Pn = None if P is None else to_host(P)
        return ll, Pn, st


class PropagatorProblem(DeviceProblem):
    """`scipy.linalg.expm(t * rate)` for batches of rate matrices `init_rate_matrix(n, v, w, pbc)` with optional
    absorbing rates on the diagonal (mcd_propagator_batch; lib/permeability/mfpt.py:364-475, perm.py:391).
    A one-lag DeviceProblem with empty counts: the kernels are the likelihood kernels, the counts are not used."""

    def __init__(self, eng: Engine, n: int, pbc: bool):
        super().__init__(eng, np.zeros((1, int(n), int(n)), dtype=np.int64), [1.0], pbc)
        if not self.squaring_ok:
            raise K.McdError("no squaring engine for %d bins" % n)

    def propagators_device(self, d_v, d_w, d_times, d_sink=None):
        """Device tensors v [B][n], w [B][dim_w], times [B], sink [B][n] or None -> P [B][n][n], status [B]."""
        eng = self.eng
        B = d_v.shape[0]
        P = torch.empty((B, self.n, self.n), dtype=torch.float64, device=eng.device)
        status = torch.empty(B, dtype=torch.int32, device=eng.device)
        K.check(eng.lib.mcd_propagator_batch(eng.handle, self.ptr, B, _dev(d_v), _dev(d_w), _dev(d_sink), _dev(d_times),
                                             _dev(P), _dev(status), eng.stream_ptr()), "mcd_propagator_batch")
        return P, status

    def propagators(self, v, w, times, sink=None):
        """Host arrays in, host arrays out.  v, w (and sink) may be one profile (broadcast over the times) or [B][..]."""
        eng = self.eng
        times = np.atleast_1d(np.asarray(times, dtype=np.float64))
        B = times.size

        def rows(a, width):
            a = np.atleast_2d(np.asarray(a, dtype=np.float64))
            if a.shape[1] != width or a.shape[0] not in (1, B):
                raise ValueError("profile shapes do not match the problem / the number of times")
            return np.ascontiguousarray(np.broadcast_to(a, (B, width)))

        with torch.cuda.device(eng.device):
            d_v = eng.to_device(rows(v, self.n), torch.float64)
            d_w = eng.to_device(rows(w, self.dim_w), torch.float64)
            d_s = None if sink is None else eng.to_device(rows(sink, self.n), torch.float64)
            d_t = eng.to_device(times, torch.float64)
            P, status = self.propagators_device(d_v, d_w, d_t, d_s)
            return to_host(P), to_host(status)


class ChainBatch:
    """State of C independent chains on one GPU (what MCState + model hold per chain)."""

    def __init__(self, prob: DeviceProblem, nchains: int, keep_basis: bool = True):
        self.prob = prob
        self.eng = prob.eng
        self.C = int(nchains)
        dev = self.eng.device
        f64, i64 = torch.float64, torch.int64
        with torch.cuda.device(dev):
            self.v = torch.zeros((self.C, prob.n), dtype=f64, device=dev)
            self.w = torch.zeros((self.C, prob.dim_w), dtype=f64, device=dev)
            self.vcoef = torch.zeros((self.C, max(prob.ncosF, 1)), dtype=f64, device=dev)
            self.wcoef = torch.zeros((self.C, max(prob.ncosD, 1)), dtype=f64, device=dev)
            self.scal = torch.zeros((self.C, K.MCD_NSCAL), dtype=f64, device=dev)
            self.counters = torch.zeros((self.C, K.MCD_NCOUNT), dtype=i64, device=dev)
            self.naccv_coeff = torch.zeros((self.C, max(prob.ncosF, 1)), dtype=i64, device=dev)
            self.naccw_coeff = torch.zeros((self.C, max(prob.ncosD, 1)), dtype=i64, device=dev)
            self.basis = None
            self.basis_valid = None
            if keep_basis and prob.in_smem:
                self.basis = torch.empty((self.C, prob.basis_elems), dtype=f64, device=dev)
                self.basis_valid = torch.zeros(self.C, dtype=torch.int32, device=dev)

    def set_state(self, v, w, vcoef=None, wcoef=None, *, dv, dw, dtimezero=0.1, temp=1.0, timezero=0.0,
                  log_like=None):
        """Upload initial profiles ([n] broadcast to all chains, or [C][n]) and MC scalars."""
        eng, prob = self.eng, self.prob

        def bc(a, width):
            a = np.asarray(a, dtype=np.float64)
            if a.ndim == 1:
                a = np.broadcast_to(a, (self.C, a.shape[0]))
            if a.shape != (self.C, width):
                raise ValueError("bad state shape %s, expected %s" % (a.shape, (self.C, width)))
            return np.ascontiguousarray(a)

        with torch.cuda.device(eng.device):
            self.v.copy_(eng.to_device(bc(v, prob.n), torch.float64))
            self.w.copy_(eng.to_device(bc(w, prob.dim_w), torch.float64))
            if prob.ncosF > 0:
                self.vcoef.copy_(eng.to_device(bc(vcoef, prob.ncosF), torch.float64))
            if prob.ncosD > 0:
                self.wcoef.copy_(eng.to_device(bc(wcoef, prob.ncosD), torch.float64))
            scal = np.zeros((self.C, K.MCD_NSCAL))
            scal[:, K.S_T0] = timezero
            scal[:, K.S_DV] = dv
            scal[:, K.S_DW] = dw
            scal[:, K.S_DT0] = dtimezero
            scal[:, K.S_TEMP] = temp
            if log_like is not None:
                scal[:, K.S_LL] = log_like
            self.scal.copy_(eng.to_device(scal, torch.float64))
            self.counters.zero_()
            self.naccv_coeff.zero_()
            self.naccw_coeff.zero_()
            if self.basis_valid is not None:
                self.basis_valid.zero_()

    def init_log_like(self):
        """First likelihood of every chain (MCState.init_log_like, lib/MCState.py:101-137).
        Returns the host copy; raises like the reference on a non-finite value."""
        prob = self.prob
        with torch.cuda.device(self.eng.device):
            lag = None
            t0 = self.scal[:, K.S_T0]
            if bool((t0 != 0).any()):
                lag = (prob.d_lag[None, :] + t0[:, None]).contiguous()
            # the engine the MC loop will use (ll_cur and the trial values must come from the same arithmetic):
            # lag times shifted by the chain's time offset have a LagPlan (plan[1]) on every squaring engine
            method = K.METHOD_SQUARING if (lag is not None and prob.squaring_ok) else K.METHOD_AUTO
            ll, _, status = prob.loglike_device(self.v, self.w, lag, method=method)
            if prob.spring_k > 0:  # lib/MCState.py:121-124
                d = self.w[:, 1:] - self.w[:, :-1]
                e = 0.5 * prob.spring_k * (d * d).sum(dim=1)
                if prob.pbc:
                    e = e + 0.5 * prob.spring_k * (self.w[:, 0] - self.w[:, -1]) ** 2
                ll = ll - e
            self.scal[:, K.S_LL] = ll
            host = to_host(ll)
        if np.any(np.isnan(host)):
            raise ValueError("Initial likelihood diverges")  # lib/MCState.py:116-117
        return host

    def run(self, nsteps, params: K.McParams, tape_u=None, tape_idx=None, record=False, nsamples_cap=0):
        """Advance all chains by nsteps (asynchronous).  Returns device tensors
        (decisions, ll_try, samples), each None unless requested."""
        eng, prob = self.eng, self.prob
        dev = eng.device
        with torch.cuda.device(dev):
            io = K.ChainIO()
            io.v, io.w = self.v.data_ptr(), self.w.data_ptr()
            io.vcoef, io.wcoef = self.vcoef.data_ptr(), self.wcoef.data_ptr()
            io.scal, io.counters = self.scal.data_ptr(), self.counters.data_ptr()
            io.naccv_coeff, io.naccw_coeff = self.naccv_coeff.data_ptr(), self.naccw_coeff.data_ptr()
            io.basis = None if self.basis is None else self.basis.data_ptr()
            io.basis_valid = None if self.basis_valid is None else self.basis_valid.data_ptr()
            keep = []
            if params.use_tape:
                tu = eng.to_device(np.asarray(tape_u, dtype=np.float64), torch.float64)
                ti = eng.to_device(np.asarray(tape_idx, dtype=np.int32), torch.int32)
                keep += [tu, ti]
                io.tape_u, io.tape_idx = tu.data_ptr(), ti.data_ptr()
            dec = llt = samples = None
            if record:
                dec = torch.zeros((self.C, nsteps), dtype=torch.uint8, device=dev)
                llt = torch.zeros((self.C, nsteps), dtype=torch.float64, device=dev)
                io.decisions, io.ll_try = dec.data_ptr(), llt.data_ptr()
            if nsamples_cap > 0:
                samples = torch.zeros((self.C, nsamples_cap, prob.sample_dim), dtype=torch.float64, device=dev)
                io.samples = samples.data_ptr()
            params.nsamples_cap = int(nsamples_cap)
            K.check(eng.lib.mcd_mc_run(eng.handle, prob.ptr, C.byref(params), C.byref(io), self.C, int(nsteps),
                                       eng.stream_ptr()), "mcd_mc_run")
            self._keep = keep
        return dec, llt, samples

    def host_state(self):
        """Device -> host copy of the whole chain state (one synchronisation)."""
        with torch.cuda.device(self.eng.device):
            out = dict(v=to_host(self.v), w=to_host(self.w), scal=to_host(self.scal),
                       counters=to_host(self.counters),
                       vcoef=to_host(self.vcoef) if self.prob.ncosF else None,
                       wcoef=to_host(self.wcoef) if self.prob.ncosD else None,
                       naccv_coeff=to_host(self.naccv_coeff), naccw_coeff=to_host(self.naccw_coeff))
        return out


class RadialProblem:
    """Device-resident radial problem (mcd_rad_problem): frozen v / w, lag times, radial count cubes, Bessel table,
    optional wrad basis.  The LagPlan and the pair-major counts are made once at creation."""

    def __init__(self, eng: Engine, v, w, pbc, lagtimes, counts, bessels, b0zeros, wrad_basis=None, clip=1e-32):
        self.eng = eng
        f64 = torch.float64
        counts = np.asarray(counts, dtype=np.float64)
        if counts.ndim != 4 or counts.shape[2] != counts.shape[3]:
            raise ValueError("counts must be [nlag][dim_rad][n][n]")
        self.nlag, self.dim_rad, self.n, _ = counts.shape
        self.lmax = int(np.asarray(bessels).shape[0])
        self.pbc = bool(pbc)
        self.ncos = 0 if wrad_basis is None else int(np.asarray(wrad_basis).shape[1])
        with torch.cuda.device(eng.device):
            keep = [eng.to_device(np.asarray(a, dtype=np.float64), f64)
                    for a in (v, w, lagtimes, counts, bessels, b0zeros)]
            d_basis = None if wrad_basis is None else eng.to_device(np.asarray(wrad_basis, dtype=np.float64), f64)
            d = K.RadProblemDesc()
            d.n, d.pbc, d.nlag, d.dim_rad, d.lmax, d.ncosDrad = self.n, int(self.pbc), self.nlag, self.dim_rad, self.lmax, self.ncos
            d.v, d.w, d.lagtimes, d.counts, d.bessels, d.b0zeros = (t.data_ptr() for t in keep)
            d.wrad_basis = None if d_basis is None else d_basis.data_ptr()
            d.clip = float(clip)
            p = C.c_void_p()
            K.check(eng.lib.mcd_rad_problem_create(eng.handle, C.byref(d), eng.stream_ptr(), C.byref(p)),
                    "mcd_rad_problem_create")      # synchronises: the staging tensors may go
            self.ptr = p
            np_, sd, nb, sq = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            K.check(eng.lib.mcd_rad_problem_info(p, C.byref(np_), C.byref(sd), C.byref(nb), C.byref(sq)))
            self.np, self.sample_dim, self.nbatch, self.squaring_ok = np_.value, sd.value, nb.value, bool(sq.value)

    def close(self):
        if getattr(self, "ptr", None):
            self.eng.lib.mcd_rad_problem_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def loglike_device(self, d_wrad, want_P=False):
        eng = self.eng
        B = d_wrad.shape[0]
        out = torch.empty(B, dtype=torch.float64, device=eng.device)
        status = torch.empty(B, dtype=torch.int32, device=eng.device)
        P = (torch.empty((B, self.nlag, self.dim_rad, self.n, self.n), dtype=torch.float64, device=eng.device)
             if want_P else None)
        K.check(eng.lib.mcd_rad_loglike_batch(eng.handle, self.ptr, B, _dev(d_wrad), _dev(out), _dev(P), _dev(status),
                                              eng.stream_ptr()), "mcd_rad_loglike_batch")
        return out, P, status


def rad_loglike(eng: Engine, v, w, pbc, lagtimes, counts, bessels, b0zeros, wrad, clip=1e-32, want_P=False):
    """Radial likelihood seam: mcdiff.twod.rad_log_like_lag (lib/twod.py:13-32)."""
    wrad = np.atleast_2d(np.asarray(wrad, dtype=np.float64))
    prob = RadialProblem(eng, v, w, pbc, lagtimes, counts, bessels, b0zeros, clip=clip)
    with torch.cuda.device(eng.device):
        out, P, status = prob.loglike_device(eng.to_device(wrad, torch.float64), want_P)
        res = to_host(out), (None if P is None else to_host(P)), to_host(status)
    prob.close()
    return res


class RadialChains:
    """C independent chains of the radial Monte-Carlo loop (only wrad is sampled; v, w frozen):
    lib/mc.py:29-30, MCState.mcmove_diffusion_radial (lib/MCState.py:306-340)."""

    def __init__(self, eng: Engine, v, w, pbc, lagtimes, counts, bessels, b0zeros, nchains, wrad_basis=None,
                 clip=1e-32):
        self.eng = eng
        f64 = torch.float64
        self.prob = RadialProblem(eng, v, w, pbc, lagtimes, counts, bessels, b0zeros, wrad_basis=wrad_basis, clip=clip)
        self.nlag, self.dim_rad, self.n, self.lmax = self.prob.nlag, self.prob.dim_rad, self.prob.n, self.prob.lmax
        self.pbc = bool(pbc)
        self.C = int(nchains)
        self.clip = float(clip)
        self.ncos = self.prob.ncos
        dev = eng.device
        with torch.cuda.device(dev):
            self.wrad = torch.zeros((self.C, self.n), dtype=f64, device=dev)
            self.wcoef = torch.zeros((self.C, max(self.ncos, 1)), dtype=f64, device=dev)
            self.scal = torch.zeros((self.C, 8), dtype=f64, device=dev)
            self.counters = torch.zeros((self.C, 8), dtype=torch.int64, device=dev)
            self.nacc_coeff = torch.zeros((self.C, max(self.ncos, 1)), dtype=torch.int64, device=dev)

    @property
    def sample_dim(self):
        return 3 + (self.ncos if self.ncos > 0 else self.n)

    def set_state(self, wrad, wcoef=None, *, dwrad, temp=1.0):
        eng = self.eng
        with torch.cuda.device(eng.device):
            wr = np.broadcast_to(np.asarray(wrad, dtype=np.float64), (self.C, self.n))
            self.wrad.copy_(eng.to_device(np.ascontiguousarray(wr), torch.float64))
            if self.ncos > 0:
                wc = np.broadcast_to(np.asarray(wcoef, dtype=np.float64), (self.C, self.ncos))
                self.wcoef.copy_(eng.to_device(np.ascontiguousarray(wc), torch.float64))
            sc = np.zeros((self.C, 8))
            sc[:, 1], sc[:, 2] = dwrad, temp
            self.scal.copy_(eng.to_device(sc, torch.float64))
            self.counters.zero_()
            self.nacc_coeff.zero_()

    def init_log_like(self):
        eng = self.eng
        with torch.cuda.device(eng.device):
            out, _, status = self.prob.loglike_device(self.wrad)
            self.scal[:, 0] = out
            host = to_host(out)
        if np.any(np.isnan(host)):
            raise ValueError("Initial likelihood diverges")
        return host

    def run(self, nsteps, *, num_mc_update, dtemp=0.0, seed=0, chain_offset=0, tape_u=None, tape_idx=None,
            record=False, nsamples_cap=0):
        eng = self.eng
        dev = eng.device
        with torch.cuda.device(dev):
            mp = K.RadMcParams()
            mp.num_mc_update, mp.dtemp, mp.sample_every = float(num_mc_update), float(dtemp), 100
            mp.use_tape, mp.ncosDrad, mp.seed = int(tape_u is not None), self.ncos, int(seed)
            mp.chain_offset, mp.tape_stride, mp.nsamples_cap = int(chain_offset), 0, int(nsamples_cap)
            io = K.RadChainIO()
            io.wrad, io.wcoef, io.scal = self.wrad.data_ptr(), self.wcoef.data_ptr(), self.scal.data_ptr()
            io.counters, io.nacc_coeff = self.counters.data_ptr(), self.nacc_coeff.data_ptr()
            io.wrad_basis = None
            keep = []
            if tape_u is not None:
                tu = eng.to_device(np.asarray(tape_u, dtype=np.float64), torch.float64)
                ti = eng.to_device(np.asarray(tape_idx, dtype=np.int32), torch.int32)
                keep += [tu, ti]
                io.tape_u, io.tape_idx = tu.data_ptr(), ti.data_ptr()
            dec = llt = smp = None
            if record:
                dec = torch.zeros((self.C, nsteps), dtype=torch.uint8, device=dev)
                llt = torch.zeros((self.C, nsteps), dtype=torch.float64, device=dev)
                io.decisions, io.ll_try = dec.data_ptr(), llt.data_ptr()
            if nsamples_cap > 0:
                smp = torch.zeros((self.C, nsamples_cap, self.sample_dim), dtype=torch.float64, device=dev)
                io.samples = smp.data_ptr()
            K.check(eng.lib.mcd_rad_mc_run(eng.handle, self.prob.ptr, C.byref(mp), C.byref(io), self.C, int(nsteps),
                                           eng.stream_ptr()), "mcd_rad_mc_run")
            self._keep = keep
        return dec, llt, smp

    def host_state(self):
        with torch.cuda.device(self.eng.device):
            return dict(wrad=to_host(self.wrad), wcoef=to_host(self.wcoef), scal=to_host(self.scal),
                        counters=to_host(self.counters), nacc_coeff=to_host(self.nacc_coeff))

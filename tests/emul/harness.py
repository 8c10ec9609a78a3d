"""ctypes driver of the CPU emulation of the CUDA chain body (tests only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mcdiff_b200 import _capi as K

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libmcd_emul.so")
SRC = os.path.join(HERE, "emul.cpp")
DEPS = [SRC] + [os.path.join(ROOT, "mcdiff_b200", "csrc", f) for f in ("mcd_chain.cuh", "mcd_exec.cuh")] + [
    os.path.join(ROOT, "include", "mcdiff_b200.h")]


def build(force=False, threads=None):
    """threads=128: the chain body as the 128-thread translation unit of the library compiles it (mcd_small.cu)."""
    so = SO if threads is None else SO.replace(".so", "%d.so" % threads)
    if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in DEPS):
        return so
    extra = [] if threads is None else ["-DMCD_THREADS=%d" % threads]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC"] + extra + [SRC, "-o", so])
    return so


def use_threads(threads=None):
    """Switch the emulator library used by loglike_batch / HostChains (None = the default 384-thread build)."""
    global _lib
    _lib = C.CDLL(build(threads=threads))
    _lib.emul_loglike_batch.restype = C.c_int
    _lib.emul_mc_run.restype = C.c_int


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emul_loglike_batch.restype = C.c_int
        _lib.emul_mc_run.restype = C.c_int
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data


class HostProblem:
    """Keeps the numpy arrays alive and exposes a ProblemDesc with host pointers."""

    def __init__(self, counts, lagtimes, pbc, v_basis=None, w_basis=None, spring_k=-1.0, string_vecs=None,
                 clip=1e-10, pull=None):
        self.counts = np.ascontiguousarray(counts, dtype=np.int32)
        self.lagtimes = np.ascontiguousarray(lagtimes, dtype=np.float64)
        self.n = self.counts.shape[-1]
        self.pbc = bool(pbc)
        self.dim_w = self.n if self.pbc else self.n - 1
        self.v_basis = None if v_basis is None else np.ascontiguousarray(v_basis, dtype=np.float64)
        self.w_basis = None if w_basis is None else np.ascontiguousarray(w_basis, dtype=np.float64)
        self.svecs = None if string_vecs is None else np.ascontiguousarray(string_vecs, dtype=np.float64)
        self.ncosF = 0 if self.v_basis is None else self.v_basis.shape[1]
        self.ncosD = 0 if self.w_basis is None else self.w_basis.shape[1]
        d = K.ProblemDesc()
        d.n, d.pbc, d.nlag = self.n, int(self.pbc), len(self.lagtimes)
        d.ncosF, d.ncosD = self.ncosF, self.ncosD
        d.lagtimes, d.counts = ptr(self.lagtimes), ptr(self.counts)
        d.v_basis, d.w_basis, d.string_vecs = ptr(self.v_basis), ptr(self.w_basis), ptr(self.svecs)
        d.spring_k, d.clip = float(spring_k), float(clip)
        d.use_pull, d.pull = (0, 0.0) if pull is None else (1, float(pull))
        self.desc = d
        self.sample_dim = 6 + (self.ncosF or self.n) + (self.ncosD or self.dim_w)


def loglike_batch(prob: HostProblem, v, w, lagtimes=None, want_P=False, method=1):
    v = np.ascontiguousarray(np.atleast_2d(v), dtype=np.float64)
    w = np.ascontiguousarray(np.atleast_2d(w), dtype=np.float64)
    B = v.shape[0]
    out = np.zeros(B)
    P = np.zeros((B, len(prob.lagtimes), prob.n, prob.n)) if want_P else None
    status = np.zeros(B, dtype=np.int32)
    sweeps = np.zeros(B, dtype=np.int32)
    lt = None if lagtimes is None else np.ascontiguousarray(lagtimes, dtype=np.float64)
    rc = lib().emul_loglike_batch(C.byref(prob.desc), C.c_int(method), C.c_int(B), C.c_void_p(ptr(v)), C.c_void_p(ptr(w)),
                                  C.c_void_p(ptr(lt)), C.c_void_p(ptr(out)), C.c_void_p(ptr(P)),
                                  C.c_void_p(ptr(status)), C.c_void_p(ptr(sweeps)))
    assert rc == 0
    return out, P, status, sweeps


def propagator_batch(prob: HostProblem, v, w, sink, times):
    """Host twin of mcd_propagator_batch: exp(t (R - diag(sink))) per item on the tensor-core squaring route of the
    chain body (a one-lag problem; per-item time and absorbing rates)."""
    assert len(prob.lagtimes) == 1
    sk = None if sink is None else np.ascontiguousarray(np.atleast_2d(sink), dtype=np.float64)
    lib().emul_set_sink(C.c_void_p(ptr(sk)))
    try:
        _, P, status, _ = loglike_batch(prob, v, w, lagtimes=np.asarray(times, dtype=np.float64)[:, None], want_P=True,
                                        method=2)
    finally:
        lib().emul_set_sink(C.c_void_p(None))
    return P[:, 0], status


class HostChains:
    """Host-memory twin of mcdiff_b200.engine.ChainState for the emulator."""

    def __init__(self, prob: HostProblem, nchains, v0, w0, vc0=None, wc0=None, ll0=0.0, dv=0.5, dw=0.5,
                 dt0=0.1, temp=1.0, timezero=0.0):
        n, dim_w = prob.n, prob.dim_w
        self.prob = prob
        self.C = nchains
        self.v = np.tile(np.asarray(v0, dtype=np.float64), (nchains, 1)).copy()
        self.w = np.tile(np.asarray(w0, dtype=np.float64), (nchains, 1)).copy()
        self.vc = np.tile(np.asarray(vc0, dtype=np.float64), (nchains, 1)).copy() if prob.ncosF else np.zeros((nchains, 1))
        self.wc = np.tile(np.asarray(wc0, dtype=np.float64), (nchains, 1)).copy() if prob.ncosD else np.zeros((nchains, 1))
        self.scal = np.zeros((nchains, K.MCD_NSCAL))
        self.scal[:, K.S_LL] = ll0
        self.scal[:, K.S_T0] = timezero
        self.scal[:, K.S_DV] = dv
        self.scal[:, K.S_DW] = dw
        self.scal[:, K.S_DT0] = dt0
        self.scal[:, K.S_TEMP] = temp
        self.counters = np.zeros((nchains, K.MCD_NCOUNT), dtype=np.int64)
        self.naccv_coeff = np.zeros((nchains, max(prob.ncosF, 1)), dtype=np.int64)
        self.naccw_coeff = np.zeros((nchains, max(prob.ncosD, 1)), dtype=np.int64)
        assert self.v.shape == (nchains, n) and self.w.shape == (nchains, dim_w)

    def run(self, nsteps, params: K.McParams, tape_u=None, tape_idx=None, record=True, nsamples_cap=0):
        io = K.ChainIO()
        io.v, io.w, io.vcoef, io.wcoef = ptr(self.v), ptr(self.w), ptr(self.vc), ptr(self.wc)
        io.scal, io.counters = ptr(self.scal), ptr(self.counters)
        io.naccv_coeff, io.naccw_coeff = ptr(self.naccv_coeff), ptr(self.naccw_coeff)
        io.basis, io.basis_valid = None, None
        keep = []
        if tape_u is not None:
            tu = np.ascontiguousarray(tape_u, dtype=np.float64)
            ti = np.ascontiguousarray(tape_idx, dtype=np.int32)
            keep += [tu, ti]
            io.tape_u, io.tape_idx = ptr(tu), ptr(ti)
        dec = np.zeros((self.C, nsteps), dtype=np.uint8) if record else None
        llt = np.zeros((self.C, nsteps)) if record else None
        samples = np.zeros((self.C, max(nsamples_cap, 1), self.prob.sample_dim)) if nsamples_cap else None
        io.decisions, io.ll_try, io.samples = ptr(dec), ptr(llt), ptr(samples)
        params.nsamples_cap = nsamples_cap
        rc = lib().emul_mc_run(C.byref(self.prob.desc), C.byref(params), C.byref(io), C.c_int(self.C),
                               C.c_int(nsteps))
        assert rc == 0
        return dec, llt, samples

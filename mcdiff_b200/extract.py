"""Trajectory -> transition count matrices on the GPU: the host mirror of `mcdiff.tools.extract`
(`lib/tools/extract.py:71-156`) and of the driver loop of `examples/create-transition-matrix/extract_Tmat.py`.

Same names and argument meaning as the reference (`transition_matrix_add1(A, x, edges, shift)`,
`count_2D(B, X, Y, Z, edges, redges, shift)`), plus the batched form the device is built for:
`transition_counts(z, edges, shifts, period)` bins a whole `[frames, particles]` trajectory once and counts every
shift in one pass (`mcd_traj_counts`).  Counting happens in libmcdiff_b200.so only -- there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _capi as K
from .engine import Engine, to_host
from .transitions import write_Tmat_cube, write_Tmat_square  # noqa: F401  (re-exported like the reference module)

_ENGINE = None


def _engine(engine=None):
    global _ENGINE
    if engine is not None:
        return engine
    if _ENGINE is None:
        _ENGINE = Engine()
    return _ENGINE


def _check_edges(edges, name="edges"):
    if isinstance(edges, torch.Tensor) and edges.is_cuda:
        # already on the device (repeated calls): the caller vouches for increasing fp64 edges
        assert edges.dim() == 1 and edges.dtype == torch.float64 and edges.is_contiguous()
        return edges
    e = np.asarray(edges, dtype=np.float64)
    if e.ndim != 1 or e.size < 1:
        raise ValueError("%s must be a one-dimensional array" % name)
    if e.size > 1 and not np.all(np.diff(e) >= 0):
        # np.digitize accepts decreasing edges too (other bin numbering); the reference never passes them
        raise ValueError("%s must be monotonically increasing" % name)
    return e


def _shifts(shifts, nframes):
    sh = np.atleast_1d(np.asarray(shifts)).astype(np.int64)
    if sh.ndim != 1 or sh.size < 1:
        raise ValueError("shifts must be a non-empty list of frame shifts")
    for s in sh:
        assert 0 < s < nframes, "shift %d: need 0 < shift < len(x) = %d (lib/tools/extract.py:73)" % (s, nframes)
    return sh


def _coords(eng, a):
    """[frames] or [frames, particles] -> contiguous fp64 device tensor [frames, particles]."""
    t = eng.to_device(a, torch.float64)
    if t.dim() == 1:
        t = t[:, None]
    if t.dim() != 2:
        raise ValueError("trajectory must be [frames] or [frames, particles]")
    return t.contiguous()


def transition_counts_device(z, edges, shifts, period=None, engine=None, out=None):
    """Counts of every shift as an int64 DEVICE tensor [len(shifts), len(edges)+1, len(edges)+1] (asynchronous).

    z: [frames] or [frames, particles] (numpy or torch, host or device); period: zpbc of the periodic fold
    `z - zpbc*floor(z/zpbc + 0.5)` or None; out: a device tensor to accumulate into."""
    eng = _engine(engine)
    e = _check_edges(edges)
    with torch.cuda.device(eng.device):
        zt = _coords(eng, z)
        T, M = zt.shape
        sh = _shifts(shifts, T)
        ne = int(e.numel() if isinstance(e, torch.Tensor) else e.size)
        H = ne + 1
        et = e if isinstance(e, torch.Tensor) else eng.to_device(e, torch.float64)
        if out is None:
            out = torch.zeros((sh.size, H, H), dtype=torch.int64, device=eng.device)
        assert out.shape == (sh.size, H, H) and out.dtype == torch.int64 and out.is_cuda and out.is_contiguous()
        sh32 = (C.c_int32 * sh.size)(*[int(s) for s in sh])
        K.check(eng.lib.mcd_traj_counts(eng.handle, C.c_void_p(zt.data_ptr()), T, M, float(period or 0.0),
                                        C.c_void_p(et.data_ptr()), ne, sh32, int(sh.size),
                                        C.c_void_p(out.data_ptr()), eng.stream_ptr()), "mcd_traj_counts")
        # zt / et must outlive the asynchronous kernels: torch frees stream-ordered on the same stream
    return out


def transition_counts(z, edges, shifts, period=None, engine=None):
    """Host int64 array [len(shifts), len(edges)+1, len(edges)+1]: counts[l, end, start]."""
    return to_host(transition_counts_device(z, edges, shifts, period, engine))


def transition_matrix_add1(A, x, edges, shift=1, engine=None):
    """`mcdiff.tools.extract.transition_matrix_add1` (lib/tools/extract.py:71-90): A[end, start] += 1 for every pair of
    frames `shift` apart of the one-dimensional trajectory x.  A has shape (len(edges)+1,)*2 and is returned."""
    x = np.asarray(x) if not isinstance(x, torch.Tensor) else x
    assert len(x.shape) == 1
    assert shift < len(x)
    assert len(A) == len(edges) + 1
    A += transition_counts(x, edges, [shift], None, engine)[0].astype(A.dtype, copy=False)
    return A


def count_2D_device(X, Y, Z, edges, redges, shifts, period=None, engine=None, out=None):
    """Radial cube counts of every shift as an int64 DEVICE tensor [len(shifts), len(redges), H, H]."""
    eng = _engine(engine)
    edges, redges = (a.cpu().numpy() if isinstance(a, torch.Tensor) else a for a in (edges, redges))
    e, re_ = _check_edges(edges), _check_edges(redges, "redges")
    with torch.cuda.device(eng.device):
        xt, yt, zt = _coords(eng, X), _coords(eng, Y), _coords(eng, Z)
        assert xt.shape == yt.shape == zt.shape
        T, M = zt.shape
        sh = _shifts(shifts, T)
        H = e.size + 1
        et, rt = eng.to_device(e, torch.float64), eng.to_device(re_, torch.float64)
        if out is None:
            out = torch.zeros((sh.size, re_.size, H, H), dtype=torch.int64, device=eng.device)
        assert out.shape == (sh.size, re_.size, H, H) and out.dtype == torch.int64 and out.is_contiguous()
        sh32 = (C.c_int32 * sh.size)(*[int(s) for s in sh])
        K.check(eng.lib.mcd_traj_counts_rad(eng.handle, C.c_void_p(xt.data_ptr()), C.c_void_p(yt.data_ptr()),
                                            C.c_void_p(zt.data_ptr()), T, M, float(period or 0.0),
                                            C.c_void_p(et.data_ptr()), int(e.size), C.c_void_p(rt.data_ptr()),
                                            int(re_.size), sh32, int(sh.size), C.c_void_p(out.data_ptr()),
                                            eng.stream_ptr()), "mcd_traj_counts_rad")
    return out


def count_2D(B, X, Y, Z, edges, redges, shift=1, engine=None):
    """`mcdiff.tools.extract.count_2D` (lib/tools/extract.py:127-156): B[r-1, end, start] += 1 with the in-plane
    displacement sqrt(dX^2 + dY^2) over `shift` frames binned in redges."""
    assert len(B.shape) == 3
    assert B.shape[0] == len(redges)
    assert B.shape[1] == len(edges) + 1
    assert B.shape[2] == len(edges) + 1
    B += to_host(count_2D_device(X, Y, Z, edges, redges, [shift], None, engine))[0].astype(B.dtype, copy=False)
    return B


def extract_Tmat(z, zpbc, nbins, shifts, outdir, dt=1.0, count="pbc", engine=None):
    """The loop of examples/create-transition-matrix/extract_Tmat.py:14-50 for a periodic box: bin the folded
    coordinates of every particle column of z [frames, particles] in `nbins` bins over [-zpbc/2, zpbc/2], count the
    transitions of every shift in one device pass and write transitions.nbins<N>.<shift>.<count>.dat files
    (interior bins only, as the example's A[1:-1, 1:-1]).  Returns the list of files."""
    dx = zpbc / nbins
    edges = np.arange(-nbins * dx / 2, (nbins + 1) * dx / 2., dx)
    assert len(edges) == nbins + 1
    A = transition_counts(z, edges, shifts, period=zpbc, engine=engine)
    os.makedirs(outdir, exist_ok=True)
    files = []
    for l, shift in enumerate(np.atleast_1d(shifts)):
        name = os.path.join(outdir, "transitions.nbins%i.%i.%s.dat" % (nbins, int(shift), count))
        write_Tmat_square(A[l, 1:-1, 1:-1], name, int(shift) * dt, count, edges=edges, dt=dt, dn=int(shift))
        files.append(name)
    return files

"""Readers of the `.dat` result files and of Logger objects.

The text layout is the reference's output contract: section markers
`----- final Settings MC -----`, `===== final v_coeff =====`, `===== final w_coeff =====`,
`===== final F D =====`, a header line starting with "   index  bin-str  bin-end", rows
`index bin-start bin-end F D [Fst Dst]`, terminator `#Done`.  Function names and return values
follow lib/outreading.py so that `--initf`, `plotresults` and `mcdiff-infty` style consumers
work unchanged.
"""
from __future__ import annotations

import numpy as np

_PROFILE_HEADER = "   index  bin-str  bin-end"


def _rows_after(filename, startline):
    """Data rows following the first line that starts with `startline`, up to a '=' line."""
    rows = []
    with open(filename) as f:
        for line in f:
            if line.startswith(startline):
                break
        for line in f:
            if line.startswith("="):
                break
            if not line.startswith("#"):
                rows.append(line.split())
    return rows


def read_F_D_edges(filename):
    """F [kBT], D [A^2/ps], edges [A] of a result file (lib/outreading.py:5-35)."""
    F, D, lo, hi = [], [], [], []
    for wds in _rows_after(filename, _PROFILE_HEADER):
        lo.append(float(wds[1]))
        hi.append(float(wds[2]))
        F.append(float(wds[3]))
        if len(wds) >= 5:
            D.append(float(wds[4]))
        elif len(wds) not in (4, 6):
            raise ValueError("error in the format line " + " ".join(wds))
    return np.array(F), np.array(D), np.array(lo + [hi[-1]])


def read_Drad(filename):
    """Radial diffusion profile (lib/outreading.py:37-64)."""
    D, lo, hi = [], [], []
    for wds in _rows_after(filename, _PROFILE_HEADER + "  diffusion-coefficient-at[i]"):
        lo.append(float(wds[1]))
        hi.append(float(wds[2]))
        if len(wds) in (4, 5):
            D.append(float(wds[3]))
    return np.array(D), np.array(lo + [hi[-1]])


def _read_coeffs(filename, tag, final):
    start = "===== final %s =====" % tag if final else "===== %s =====" % tag
    return np.array([float(wds[1]) for wds in _rows_after(filename, start)])


def read_Fcoeffs(filename, final=False):
    return _read_coeffs(filename, "v_coeff", final)


def read_Dcoeffs(filename, final=False):
    """First coefficient is stored shifted by +wunit (lib/log.py:298)."""
    return _read_coeffs(filename, "w_coeff", final)


def read_Dradcoeffs(filename, final=False):
    return _read_coeffs(filename, "wrad_coeff", final)


def read_dv_dw(filename, final=False):
    start = "----- final Settings MC -----" if final else "----- Settings MC -----"
    dv = dw = None
    with open(filename) as f:
        for line in f:
            if line.startswith(start):
                break
        for line in f:
            if line.startswith("dv"):
                dv = float(line.split()[1])
                break
        for line in f:
            if line.startswith("dw"):
                dw = float(line.split()[1])
                break
    return dv, dw


def read_many_profiles(list_filename, pic=False):
    """F (shifted to min 0), D, edges of several result files (lib/outreading.py:66-97)."""
    F, D, E = [], [], []
    if pic:
        from .log import load_logger
        Fst, Dst = [], []
        for fn in list_filename:
            f, d, e, fst, dst = read_F_D_edges_logger(load_logger(fn))
            F.append(f - min(f)); D.append(d); E.append(e); Fst.append(fst); Dst.append(dst)
        return F, D, E, Fst, Dst
    for fn in list_filename:
        f, d, e = read_F_D_edges(fn)
        F.append(f - min(f)); D.append(d); E.append(e)
    return F, D, E, None, None


# ---- Logger statistics (all stored rows, including the initial one: lib/outreading.py:281-301) ----
def average_profile(samples):
    samples = np.asarray(samples)
    return np.mean(samples, axis=0), np.std(samples, axis=0)


def read_coeff_logger(logger):
    m = logger.model
    v_coeff = v_st = w_coeff = w_st = wrad_coeff = wrad_st = None
    if m.ncosF > 0:
        v_coeff, v_st = average_profile(logger.v_coeff)
    if m.ncosD > 0:
        w_coeff, w_st = average_profile(logger.w_coeff)
    if getattr(m, "ncosDrad", 0) > 0:
        wrad_coeff, wrad_st = average_profile(logger.wrad_coeff)
    return (v_coeff, w_coeff, wrad_coeff, v_st, w_st, wrad_st, np.mean(logger.timezero), np.std(logger.timezero))


def read_F_D_edges_logger(logger):
    """Posterior mean/std of F and of D = exp(w) (mean of exp, not exp of mean)."""
    m = logger.model
    if m.ncosF <= 0:
        v, vst = average_profile(logger.v)
    else:
        v, vst = logger.average_profile_v_from_coeff()
    if m.ncosD <= 0:
        d, dst = average_profile(np.exp(logger.w))
    else:
        _, _, d, dst = logger.average_profile_w_from_coeff(want_log=False)
    unit = np.exp(m.wunit)
    return v * m.vunit, d * unit, m.edges, vst * m.vunit, dst * unit


def read_Drad_logger(logger):
    m = logger.model
    if not hasattr(m, "wradunit"):
        return None, None, None
    if getattr(m, "ncosDrad", 0) <= 0:
        drad, dradst = average_profile(np.exp(logger.wrad))
    else:
        _, _, drad, dradst = logger.average_profile_wrad_from_coeff(want_log=False)
    unit = np.exp(m.wradunit)
    return drad * unit, m.redges, dradst * unit

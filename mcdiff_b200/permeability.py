"""Propagator users of the permeability analyses on the GPU: the `scipy.linalg.expm(rate*t)`
call sites of `lib/permeability/mfpt.py:364,373,475`, `lib/permeability/perm.py:391` and `lib/MCState.py:353`, served by
one batched device call (`mcd_propagator_batch`: every time of a grid is one CTA).

Host mirror of the reference's argument conventions: F in kBT, D in A^2/ps, dx in A, dt in ps, rates in 1/dt
(`lib/utils.py:186-222`), absorbing bins st / end and `side` as `init_rate_matrix` takes them (`lib/utils.py:28-74`).
No plotting (the reference's matplotlib figures are out of scope); the numbers behind the figures are returned.
There is no CPU path: the matrix exponentials come from libmcdiff_b200.so only.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine, PropagatorProblem

_ENGINE = None
_PROBLEMS = {}


def _problem(n, pbc, engine=None):
    global _ENGINE
    if engine is None:
        if _ENGINE is None:
            _ENGINE = Engine()
        engine = _ENGINE
    key = (id(engine), int(n), bool(pbc))
    if key not in _PROBLEMS:
        _PROBLEMS[key] = PropagatorProblem(engine, n, pbc)
    return _PROBLEMS[key]


def profiles_from_F_D(F, D, dx, dt):
    """v, w of `construct_rate_matrix_from_F_D` (lib/utils.py:209-212): w = ln D - ln(dx^2/dt), one value per bin."""
    F = np.asarray(F, dtype=np.float64)
    w = np.log(np.asarray(D, dtype=np.float64)) - np.log(dx ** 2 / dt)
    return F, w


def absorbing_cut(v, w, st=None, end=None, side=None):
    """The cut rate matrix of `init_rate_matrix(n, v, w, False, st=, end=, side=)` (lib/utils.py:56-73) as the pieces the
    device needs: (v_sub, w_sub, sink) with v_sub = v[st+1:end], w_sub = the len-1 couplings between those bins and
    sink = rates from the first / last kept bin into the removed neighbours st / end (periodic neighbours, because the
    reference cuts the PERIODIC matrix; zero on a side that is made reflective)."""
    v, w = np.asarray(v, dtype=np.float64), np.asarray(w, dtype=np.float64)
    n = len(v)
    assert len(w) == n, "the cut is taken from the periodic matrix: w has one value per bin (lib/utils.py:59)"
    if st is None:
        st = -1
    if end is None:
        end = n
    assert st >= -1 and st < n - 1
    assert end <= n and end > 1
    assert st + 1 < end
    lo, hi = st + 1, end - 1                       # first and last kept bin
    m = hi - lo + 1
    sink = np.zeros(m)
    # The diagonal of the periodic matrix keeps BOTH losses of every bin, also after the corner elements are zeroed
    # (lib/utils.py:59-61) and the rows/columns of the neighbours are cut away: what the first / last kept bin loses
    # to its removed (periodic) neighbour is absorbed.  A side that is not absorbing gets the reflective diagonal
    # (lib/utils.py:70-73), i.e. no sink.
    if side in ("left", "both"):
        j = (lo - 1) % n                           # rate lo -> lo-1: exp(w[lo-1] + (v[lo] - v[lo-1])/2)
        sink[0] = np.exp(w[j] - 0.5 * (v[j] - v[lo]))
    if side in ("right", "both"):
        j = (hi + 1) % n                           # rate hi -> hi+1: exp(w[hi] - (v[hi+1] - v[hi])/2)
        sink[-1] += np.exp(w[hi] - 0.5 * (v[j] - v[hi]))
    return v[lo:hi + 1].copy(), w[lo:hi].copy(), sink


def propagators(F, D, dx, dt, times, pbc=True, st=None, end=None, side=None, engine=None):
    """exp(t * rate) for every t of `times`, rate = construct_rate_matrix_from_F_D(F, D, dx, dt, pbc, st=, end=, side=)
    (lib/utils.py:202-222).  Returns [len(times), m, m]."""
    v, w = profiles_from_F_D(F, D, dx, dt)
    times = np.atleast_1d(np.asarray(times, dtype=np.float64))
    if pbc:
        if st is not None or end is not None or side is not None:
            raise NotImplementedError("pbc=True with absorption/reflection (lib/utils.py:39-43)")
        P, status = _problem(len(v), True, engine).propagators(v, w, times)
    elif st is None and end is None and side is None:
        P, status = _problem(len(v), False, engine).propagators(v, w[:-1], times)      # lib/utils.py:217
    else:
        vs, ws, sink = absorbing_cut(v, w, st, end, side)
        P, status = _problem(len(vs), False, engine).propagators(vs, ws, times, sink=sink)
    if status.any():
        raise FloatingPointError("non-finite rate matrix")
    return P


def distribution_mfpt(F, D, dx, dt, b1, b2, trange, filename=None, engine=None):
    """`mcdiff.permeability.mfpt.distribution_mfpt` (lib/permeability/mfpt.py:449-492): exit-time density rho(t | k)
    and its cumulative distribution for the middle bin k of the interval between the absorbing bins b1 and b2, on the
    time grid `trange` -- one batched propagator call instead of a Python loop of expm's.  Returns (rho, cdf); writes
    the three-column file the reference calls t_rho.dat when `filename` is given."""
    assert b1 >= -1 and b2 <= len(F) and b1 + 1 < b2
    F = np.asarray(F, dtype=np.float64)
    v, w = profiles_from_F_D(F - np.min(F), D, dx, dt)
    vs, ws, sink = absorbing_cut(v, w, b1, b2, "both")
    m = len(vs)
    trange = np.asarray(trange, dtype=np.float64)
    P, status = _problem(m, False, engine).propagators(vs, ws, trange, sink=sink)
    if status.any():
        raise FloatingPointError("non-finite rate matrix")
    middle = m // 2                                 # len(subrate)/2 under Python 2 division (mfpt.py:468)
    # rho_t = -sum(dot(subrate, mat), 0): the column sums of subrate are -sink, so rho_t[k] = sum_i sink_i mat[i, k]
    rho = np.einsum("i,tik->tk", sink, P)[:, middle]
    cdf = (1.0 - P.sum(axis=1))[:, middle]          # sum(I - mat, 0)
    if filename is not None:
        np.savetxt(filename, np.column_stack([trange, rho, cdf]))
    return rho, cdf


def partitioning_transient(F, D, pbc, dt, st, end, times, p_0, engine=None):
    """The table behind `analyze_partitioning_transient` (lib/permeability/perm.py:385-405): for every lag time the
    concentration profile prop . p_0 and, from it, [p_mem, p_wat, prob[st], prob[end], flux[st], flux[end-1], p_wat2].
    As the reference, w = ln D is used as it is (no unit shift).  Returns (data [len(times), 7], prob [len(times), n])."""
    F, D = np.asarray(F, dtype=np.float64), np.asarray(D, dtype=np.float64)
    n = len(F)
    w = np.log(D)
    P, status = _problem(n, pbc, engine).propagators(F, w if pbc else w[:n - 1], times)
    if status.any():
        raise FloatingPointError("non-finite rate matrix")
    prob = P @ np.asarray(p_0, dtype=np.float64)
    up = np.exp(w[:n - 1] - 0.5 * (F[1:] - F[:-1]))          # rate[i+1, i]   (lib/utils.py:96)
    down = np.exp(w[:n - 1] + 0.5 * (F[1:] - F[:-1]))        # rate[i, i+1]   (lib/utils.py:97)
    flux = (up * prob[:, :-1] - down * prob[:, 1:]) / dt     # calc_flux, perm.py:128-148
    seg = prob[:, st:end + 1]
    data = np.zeros((len(prob), 7))
    data[:, 0] = seg[:, 1:-1].sum(axis=1) / (seg.shape[1] - 2)          # calc_mean_prob, perm.py:282-286
    data[:, 1] = (seg[:, 0] + seg[:, -1]) / 2.0
    data[:, 2] = prob[:, st]
    data[:, 3] = prob[:, end]
    data[:, 4] = flux[:, st]
    data[:, 5] = flux[:, end - 1]
    data[:, 6] = (prob[:, :st + 1].sum(axis=1) + prob[:, end:].sum(axis=1)) / (n - end + st + 1)
    return data, prob

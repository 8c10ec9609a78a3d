"""CPU oracle for the propagator users of the permeability analyses.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/`` may import it.

Restates in numpy / scipy:

* ``init_rate_matrix`` with absorbing / reflective cuts  -- /root/reference/lib/utils.py:28-74 (+ :77-131)
* ``construct_rate_matrix_from_F_D``                    -- /root/reference/lib/utils.py:202-222
* ``distribution_mfpt`` (rho, cdf on a time grid)       -- /root/reference/lib/permeability/mfpt.py:449-492
* the table of ``analyze_partitioning_transient``        -- /root/reference/lib/permeability/perm.py:385-405,
  ``calc_flux`` :128-148, ``calc_mean_prob`` :282-286

The matrix exponential is ``scipy.linalg.expm`` (third party, see oracle/mcdiff_oracle.py).  Parity pinning: the rate
matrices and propagators are checked against ``tests/golden/perm_golden.npz``, produced by calling the unmodified
reference's ``construct_rate_matrix_from_F_D`` / ``init_rate_matrix`` and applying the formulas of the call sites
(``lib/permeability/mfpt.py`` itself needs matplotlib and Python-2 integer division, so the module cannot be imported
here; ``tests/golden/make_perm_golden.py`` restates only its three arithmetic lines :475-485 on the reference's matrix).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm

from oracle import mcdiff_oracle as O


def init_rate_matrix(n, v, w, pbc, st=None, end=None, side=None):
    """lib/utils.py:28-74 without the pull branch."""
    if pbc:
        if st is not None or end is not None or side is not None:
            raise NotImplementedError
        return O.rate_matrix(v, w, True)
    if st is None and end is None and side is None:
        return O.rate_matrix(v, w, False)
    rate = O.rate_matrix(v, w, True)
    rate[-1, 0] = 0.0
    rate[0, -1] = 0.0
    if st is None:
        st = -1
    if end is None:
        end = n
    assert -1 <= st < n - 1 and 1 < end <= n and st + 1 < end
    rate = rate[st + 1:end, st + 1:end].copy()
    if side not in ("left", "both"):
        rate[0, 0] = -rate[1, 0]
    if side not in ("right", "both"):
        rate[-1, -1] = -rate[-2, -1]
    return rate


def construct_rate_matrix_from_F_D(F, D, dx, dt, pbc=True, st=None, end=None, side=None):
    """lib/utils.py:202-222."""
    n = len(F)
    w = np.log(D) - np.log(dx ** 2 / dt)
    if pbc or st is not None or end is not None or side is not None:
        return init_rate_matrix(n, np.asarray(F, float), w, pbc, st=st, end=end, side=side)
    return init_rate_matrix(n, np.asarray(F, float), w[:-1], False)


def distribution_mfpt(F, D, dx, dt, b1, b2, trange):
    """lib/permeability/mfpt.py:449-485 (numbers only)."""
    F = np.asarray(F, float)
    subrate = construct_rate_matrix_from_F_D(F - F.min(), D, dx, dt, pbc=False, st=b1, end=b2, side="both")
    middle = len(subrate) // 2
    rho, cdf = [], []
    for t in trange:
        mat = expm(subrate * t)
        rho.append((-np.sum(np.dot(subrate, mat), 0))[middle])
        cdf.append(np.sum(np.diag(np.ones(len(subrate))) - mat, 0)[middle])
    return np.array(rho), np.array(cdf)


def partitioning_transient(F, D, pbc, dt, st, end, times, p_0):
    """lib/permeability/perm.py:385-405."""
    F, D = np.asarray(F, float), np.asarray(D, float)
    n = len(F)
    w = np.log(D)
    rate = init_rate_matrix(n, F, w if pbc else w[:-1], pbc)
    data = np.zeros((len(times), 7))
    for i, t in enumerate(times):
        prob = expm(t * rate) @ p_0
        seg = prob[st:end + 1]
        fluxp = np.array([rate[k + 1, k] * prob[k] for k in range(n - 1)]) / dt
        fluxn = np.array([rate[k, k + 1] * prob[k + 1] for k in range(n - 1)]) / dt
        flux = fluxp - fluxn
        data[i] = [np.sum(seg[1:-1]) / (len(seg) - 2), (seg[0] + seg[-1]) / 2.0, prob[st], prob[end], flux[st],
                   flux[end - 1], (np.sum(prob[:st + 1]) + np.sum(prob[end:])) / (n - end + st + 1)]
    return data

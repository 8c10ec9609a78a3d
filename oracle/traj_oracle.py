"""CPU oracle for trajectory -> transition counts.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/`` and ``bench.py``'s CPU-baseline leg may import it;
``mcdiff_b200.extract`` never falls back to anything in this directory.

It restates, in vectorised numpy (``np.digitize`` + ``np.add.at`` instead of the reference's Python ``zip`` loops):

* ``transition_matrix_add1``  -- /root/reference/lib/tools/extract.py:71-90   (A[end, start] += 1)
* ``count_2D``                -- /root/reference/lib/tools/extract.py:127-156 (B[r-1, end, start] += 1, r = 0 wraps)
* the periodic fold and the loop over particles / shifts
                              -- /root/reference/examples/create-transition-matrix/extract_Tmat.py:36-50

Parity pinning: checked against ``tests/golden/traj_golden.npz``, produced by calling the unmodified reference
functions (``tests/golden/make_traj_golden.py``).  Integer work: results must be bit-identical.
"""
from __future__ import annotations

import numpy as np


def fold(z, zpbc):
    """extract_Tmat.py:45 -- put z = 0 in the centre of the periodic box."""
    z = np.asarray(z, dtype=np.float64)
    return z - zpbc * np.floor(z / zpbc + 0.5)


def transition_matrix_add1(A, x, edges, shift=1):
    """extract.py:71-90 for one 1-D trajectory (same asserts), vectorised."""
    x = np.asarray(x)
    assert len(x.shape) == 1
    assert shift < len(x)
    assert len(A) == len(edges) + 1
    digitized = np.digitize(x, edges)
    np.add.at(A, (digitized[shift:], digitized[:-shift]), 1)
    return A


def transition_counts(z, edges, shifts, period=None):
    """All particles (columns of z [frames, particles]) and all shifts: int64 [len(shifts), H, H], H = len(edges)+1."""
    z = np.asarray(z, dtype=np.float64)
    if z.ndim == 1:
        z = z[:, None]
    if period:
        z = fold(z, period)
    H = len(edges) + 1
    out = np.zeros((len(shifts), H, H), dtype=np.int64)
    dig = np.digitize(z, edges)
    for l, shift in enumerate(shifts):
        assert 0 < shift < z.shape[0]
        np.add.at(out[l], (dig[shift:].ravel(), dig[:-shift].ravel()), 1)
    return out


def count_2D(B, X, Y, Z, edges, redges, shift=1):
    """extract.py:127-156 for one particle (1-D X, Y, Z) or many (columns), vectorised."""
    assert len(B.shape) == 3
    assert B.shape[0] == len(redges)
    assert B.shape[1] == len(edges) + 1
    assert B.shape[2] == len(edges) + 1
    X, Y, Z = (np.asarray(a, dtype=np.float64) for a in (X, Y, Z))
    digitized = np.digitize(Z, edges)
    zstart, zend = digitized[:-shift], digitized[shift:]
    dX = X[shift:] - X[:-shift]
    dY = Y[shift:] - Y[:-shift]
    dR = np.sqrt(dX ** 2 + dY ** 2)
    dr = np.digitize(dR, redges)
    # B[r-1, end, start] += 1: r = 0 addresses B[-1], the last radial bin, exactly as the reference's indexing does
    np.add.at(B, ((dr - 1).ravel() % B.shape[0], zend.ravel(), zstart.ravel()), 1)
    return B

"""Host mirror of the device RNG: Philox4x32-10 keyed by the run seed, counter
(step_lo, step_hi, chain, stream).  Used to reproduce, on the host, exactly the proposals
and uniforms a chain consumed (mcd_chain.cuh: philox4x32_10 / u53)."""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def u53(hi, lo):
    return ((hi >> 5) * 67108864.0 + (lo >> 6)) / 9007199254740992.0


def chain_tape(seed, chain, step0, nsteps, n, dim_w, ncosF, ncosD, dv_positive=True, dw_positive=True):
    """tape_u [nsteps][5], tape_idx [nsteps] that the device derives for `chain` (global id)
    at steps step0 .. step0+nsteps-1.  The index range follows the move type exactly as the
    kernel does (potential: [0,n) or [1,ncosF); diffusion: [0,dim_w) or [0,ncosD))."""
    k0, k1 = seed & MASK, (seed >> 32) & MASK
    u = np.zeros((nsteps, 5))
    idx = np.zeros(nsteps, dtype=np.int32)
    for s in range(nsteps):
        imc = step0 + s
        lo32, hi32 = imc & MASK, (imc >> 32) & MASK
        r0 = philox4x32_10(lo32, hi32, chain & MASK, 0, k0, k1)
        r1 = philox4x32_10(lo32, hi32, chain & MASK, 1, k0, k1)
        r2 = philox4x32_10(lo32, hi32, chain & MASK, 2, k0, k1)
        u[s] = (u53(r0[0], r0[1]), u53(r0[2], r0[3]), u53(r1[2], r1[3]), u53(r2[0], r2[1]), u53(r2[2], r2[3]))
        if u[s, 0] < 0.5 and ncosF != 1 and dv_positive:
            lo, rng = (0, n) if ncosF <= 0 else (1, ncosF - 1)
        elif dw_positive:
            lo, rng = (0, dim_w) if ncosD <= 0 else (0, ncosD)
        else:
            lo, rng = 0, 1
        idx[s] = lo + ((r1[0] * rng) >> 32)
    return u, idx

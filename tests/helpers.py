"""Shared helpers for the parity tests."""
import numpy as np

REPLAY_TAGS = ["cosA", "binP4", "cosP4t0", "binCut", "springP4"]
PULL_TAGS = ["cosPull", "binPull"]     # tests/golden/pull_golden.npz (--pull trajectories of the reference)


def replay_case(R, tag):
    """Unpack one reference trajectory fixture into plain python values."""
    g = lambda k: R[tag + "/" + k]  # noqa: E731
    from oracle import mcdiff_oracle as O
    pbc = bool(g("cfg_pbc"))
    counts = g("counts")
    n = counts.shape[-1]
    dim_w = n if pbc else n - 1
    ncosF, ncosD = int(g("cfg_ncosF")), int(g("cfg_ncosD"))
    k = float(g("cfg_k"))
    te = float(g("cfg_temp_end"))
    nmc = int(g("cfg_nmc"))
    nupd = float(g("cfg_nmc_update"))
    temp = float(g("cfg_temp"))
    nupdate = nmc / nupd - 1
    dtemp = 0.0 if (np.isnan(te) or nupdate <= 0) else (te - temp) / nupdate
    return dict(
        g=g, pbc=pbc, counts=counts, lagtimes=g("lagtimes"), n=n, dim_w=dim_w, ncosF=ncosF, ncosD=ncosD, k=k,
        v_basis=O.cos_basis_center(n, ncosF) if ncosF else None,
        w_basis=O.cos_basis_border(dim_w, n, ncosD) if ncosD else None,
        svecs=O.string_vecs(dim_w, pbc) if k > 0 else None,
        nmc=nmc, nmc_update=nupd, temp=temp, temp_end=None if np.isnan(te) else te, dtemp=dtemp,
        dv=float(g("cfg_dv")), dw=float(g("cfg_dw")), dt0=float(g("cfg_dtimezero")),
        move_timezero=bool(g("cfg_move_timezero")),
        pull=float(g("dF")) if (tag + "/dF") in R.files else None)     # dF per bin (lib/MCState.py:65-67)

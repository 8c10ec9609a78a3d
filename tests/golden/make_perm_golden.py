"""Golden vectors for the permeability propagators, produced with the UNMODIFIED reference (build container only):

    python tests/golden/make_perm_golden.py        # needs /root/reference, writes tests/golden/perm_golden.npz

Rate matrices come from mcdiff.utils.construct_rate_matrix_from_F_D / init_rate_matrix (lib/utils.py:28-74,202-222).
lib/permeability/mfpt.py cannot be imported here (matplotlib; Python-2 division at :468), so the three arithmetic lines
of distribution_mfpt (:475-485) and of analyze_partitioning_transient (perm.py:390-405) are applied to the reference's
own matrices with scipy.linalg.expm, exactly as those call sites do."""
import os
import sys
import tempfile

import numpy as np
import scipy.linalg

HERE = os.path.dirname(os.path.abspath(__file__))
tmp = tempfile.mkdtemp()
os.symlink("/root/reference/lib", os.path.join(tmp, "mcdiff"))
sys.path.insert(0, tmp)
from mcdiff.utils import construct_rate_matrix_from_F_D, init_rate_matrix  # noqa: E402

rng = np.random.default_rng(77)
n, dx, dt = 60, 0.7, 1.0
x = (np.arange(n) + 0.5) / n
F = 3.0 * np.exp(-((x - 0.5) / 0.12) ** 2) + 0.3 * rng.normal(size=n)
D = 0.4 + 0.25 * np.cos(2 * np.pi * x) + 0.02 * rng.random(n)
out = dict(F=F, D=D, dx=dx, dt=dt)

# (1) cut matrices for the combinations of st / end / side the reference supports
cuts = [(9, 50, "both"), (9, 50, "left"), (9, 50, "right"), (-1, 40, "both"), (20, n, "both"), (-1, n, "both"),
        (5, 37, None)]
for k, (st, end, side) in enumerate(cuts):
    v = F - F.min()
    R = construct_rate_matrix_from_F_D(v, D, dx, dt, pbc=False, st=st, end=end, side=side)
    out["cut%d_args" % k] = np.array([st, end, {"both": 0, "left": 1, "right": 2, None: 3}[side]])
    out["cut%d_rate" % k] = R
    out["cut%d_prop" % k] = np.array([scipy.linalg.expm(R * t) for t in (0.5, 20.0, 300.0)])
out["cut_times"] = np.array([0.5, 20.0, 300.0])
out["ncuts"] = len(cuts)

# (2) distribution_mfpt, mfpt.py:459-485
b1, b2 = 9, 50
trange = np.concatenate([np.arange(1.0, 20.0, 3.0), np.logspace(1.5, 3.5, 12)])
subrate = construct_rate_matrix_from_F_D(F - F.min(), D, dx, dt, pbc=False, st=b1, end=b2, side="both")
middle = len(subrate) // 2
rho, cdf = [], []
for t in trange:
    mat = scipy.linalg.expm(subrate * t)
    rho_t = -np.sum(np.dot(subrate, mat), 0)
    rho.append(rho_t[middle])
    cdf_t = np.diag(np.ones(len(subrate))) - mat
    cdf.append(np.sum(cdf_t, 0)[middle])
out.update(mfpt_b=np.array([b1, b2]), mfpt_trange=trange, mfpt_rho=np.array(rho), mfpt_cdf=np.array(cdf))

# (3) analyze_partitioning_transient, perm.py:385-405 (pbc and reflecting)
times = np.array([1.0, 5.0, 25.0, 125.0, 625.0])
st, end = 15, 44
p_0 = np.zeros(n)
p_0[:5] = 1.0
p_0[-5:] = 1.0
for tag, pbc in (("pbc", True), ("ref", False)):
    data = np.zeros((len(times), 7))
    probs = []
    for i, lagtime in enumerate(times):
        rate = init_rate_matrix(n, F, np.log(D) if pbc else np.log(D)[:-1], pbc)
        prop = scipy.linalg.expm(lagtime * rate)
        prob = np.dot(prop, p_0)
        probs.append(prob)
        seg = prob[st:end + 1]
        p_mem, p_wat = np.sum(seg[1:-1]) / (len(seg) - 2), (seg[0] + seg[-1]) / 2.
        p_wat2 = (np.sum(prob[:st + 1]) + np.sum(prob[end:])) / (len(F) - end + st + 1)
        fluxp = np.array([rate[k + 1, k] * prob[k] for k in range(n - 1)]) / dt
        fluxn = np.array([rate[k, k + 1] * prob[k + 1] for k in range(n - 1)]) / dt
        flux = fluxp - fluxn
        data[i] = [p_mem, p_wat, prob[st], prob[end], flux[st], flux[end - 1], p_wat2]
    out["part_%s_data" % tag] = data
    out["part_%s_prob" % tag] = np.array(probs)
out.update(part_times=times, part_st_end=np.array([st, end]), part_p0=p_0)

np.savez_compressed(os.path.join(HERE, "perm_golden.npz"), **out)
print({k: np.shape(v) for k, v in out.items()})

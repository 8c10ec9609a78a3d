"""Golden vectors for trajectory -> counts, produced by the UNMODIFIED reference (run in the build container only):

    python tests/golden/make_traj_golden.py        # needs /root/reference, writes tests/golden/traj_golden.npz

Calls mcdiff.tools.extract.transition_matrix_add1 / count_2D (lib/tools/extract.py:71-90, 127-156) the way
examples/create-transition-matrix/extract_Tmat.py:36-50 does: per particle, per shift, on folded coordinates."""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
tmp = tempfile.mkdtemp()
os.symlink("/root/reference/lib", os.path.join(tmp, "mcdiff"))
sys.path.insert(0, tmp)
from mcdiff.tools.extract import count_2D, transition_matrix_add1  # noqa: E402

rng = np.random.default_rng(20261020)
out = {}

# (1) periodic box, uniform edges as the example builds them, 7 particles, random walks that leave the box
zpbc, nbins, T, M = 67.92547, 100, 400, 7
dx = zpbc / nbins
edges = np.arange(-nbins * dx / 2, (nbins + 1) * dx / 2., dx)
z = np.cumsum(rng.normal(0.0, 1.3, size=(T, M)), axis=0) + rng.uniform(-40, 40, size=M)
shifts = [1, 2, 5, 10, 50]
A_all = []
with contextlib.redirect_stdout(io.StringIO()):
    for shift in shifts:
        A = np.zeros((nbins + 2, nbins + 2), int)
        for i in range(M):
            traj = z[:, i] - zpbc * np.floor(z[:, i] / zpbc + 0.5)
            A = transition_matrix_add1(A, traj, edges, shift=shift)
        A_all.append(A)
out.update(pbc_z=z, pbc_zpbc=zpbc, pbc_edges=edges, pbc_shifts=np.array(shifts), pbc_counts=np.array(A_all))

# (2) no fold, NON-uniform edges, values outside the edges and exactly ON edges, one particle
edges2 = np.array([-3.0, -2.5, -1.0, -0.25, 0.0, 0.125, 1.0, 1.5, 4.0])
x2 = rng.normal(0.0, 2.0, size=300)
x2[::17] = edges2[rng.integers(0, len(edges2), size=len(x2[::17]))]      # hits on edges: bin = number of edges <= x
x2[5], x2[6] = -100.0, 100.0
shifts2 = [1, 3, 299]
A2 = []
with contextlib.redirect_stdout(io.StringIO()):
    for shift in shifts2:
        A2.append(transition_matrix_add1(np.zeros((len(edges2) + 1,) * 2, int), x2, edges2, shift=shift))
out.update(nu_x=x2, nu_edges=edges2, nu_shifts=np.array(shifts2), nu_counts=np.array(A2))

# (3) radial cube: 3 particles, displacement bins incl. dR beyond the last radial edge
edges3 = np.linspace(-8.0, 8.0, 17)
redges3 = np.arange(0.0, 6.1, 0.5)
T3, M3 = 250, 3
X = np.cumsum(rng.normal(0, 0.6, size=(T3, M3)), axis=0)
Y = np.cumsum(rng.normal(0, 0.6, size=(T3, M3)), axis=0)
Z = np.cumsum(rng.normal(0, 0.8, size=(T3, M3)), axis=0)
shifts3 = [1, 4, 20]
B_all = []
with contextlib.redirect_stdout(io.StringIO()):
    for shift in shifts3:
        B = np.zeros((len(redges3), len(edges3) + 1, len(edges3) + 1), int)
        for i in range(M3):
            B = count_2D(B, X[:, i], Y[:, i], Z[:, i], edges3, redges3, shift=shift)
        B_all.append(B)
out.update(rad_X=X, rad_Y=Y, rad_Z=Z, rad_edges=edges3, rad_redges=redges3, rad_shifts=np.array(shifts3),
           rad_counts=np.array(B_all))

np.savez_compressed(os.path.join(HERE, "traj_golden.npz"), **out)
print({k: np.shape(v) for k, v in out.items()})

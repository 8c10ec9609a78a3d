set -e
cd $GRAFT_REPO_ROOT
python - <<'PY'
import numpy as np, os
from mcdiff_b200.transitions import write_Tmat_square
G = np.load("tests/golden/pull_golden.npz")
os.makedirs("gpurun_out/clitmp", exist_ok=True)
for l, lag in enumerate(G["lags"]):
    write_Tmat_square(G["counts"][l], "gpurun_out/clitmp/transitions.nbins100.%d.pbc.dat" % int(lag), float(lag), "pbc", edges=G["edges"], dt=1.0, dn=int(lag))
PY
python scripts/run-mcdiff --model CosinusModel --ncosF 6 --ncosD 4 -n 2000 --nmc_update 500 -T 1 --Tend 1 --dv 0.3 --dw 0.3 -s 3 --pull 0.4 --nchains 64 -o gpurun_out/clitmp/out.pull.dat gpurun_out/clitmp/transitions.nbins100.10.pbc.dat gpurun_out/clitmp/transitions.nbins100.20.pbc.dat gpurun_out/clitmp/transitions.nbins100.30.pbc.dat gpurun_out/clitmp/transitions.nbins100.40.pbc.dat 2>&1 | tail -12
ls gpurun_out/clitmp | head

"""Multi-GPU path on real devices (skipped with fewer than 2 GPUs): chains sharded over two
ranks by torchrun give exactly the trajectories of a single-GPU run (Philox streams are keyed
by the global chain index); the only collective is the final NCCL all_gather."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_equal_one_rank(engine, gold_ll, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from mcdiff_b200.transitions import write_Tmat_square
    G = gold_ll
    tmp = str(tmp_path)
    files = []
    for l, lag in enumerate(G["P_lags"]):
        fn = os.path.join(tmp, "transitions.nbins100.%d.pbc.dat" % int(lag))
        write_Tmat_square(G["P_counts"][l], fn, float(lag), "pbc", edges=G["P_edges"], dt=1.0, dn=int(lag))
        files.append(fn)
    common = ["--model", "CosinusModel", "--ncosF", "6", "--ncosD", "4", "-n", "200", "--nmc_update", "100", "-T", "1",
              "--Tend", "1", "--dv", "0.1", "--dw", "0.1", "-s", "11", "--nchains", "7"]
    script = os.path.join(ROOT, "scripts", "run-mcdiff")
    out1 = os.path.join(tmp, "one.dat")
    out2 = os.path.join(tmp, "two.dat")
    env = dict(os.environ)
    r1 = subprocess.run([sys.executable, script, "-o", out1] + common + files, capture_output=True, text=True, env=env)
    assert r1.returncode == 0, r1.stderr[-2000:]
    r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                         "--master-addr", "127.0.0.1", "--master-port", "29533", script, "-o", out2] + common + files,
                        capture_output=True, text=True, env=env)
    assert r2.returncode == 0, r2.stderr[-2000:]
    a, b = np.load(out1 + ".chains.npz"), np.load(out2 + ".chains.npz")
    assert a["log_like"].shape == (7,) and b["log_like"].shape == (7,)
    for key in ("log_like", "naccv", "naccw", "v_param", "w_param", "dv", "dw"):
        np.testing.assert_array_equal(a[key], b[key])
    assert open(out1).read() == open(out2).read()          # chain 0 file identical

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
    # pooled posterior: rank 0 holds the Logger rows of ALL chains of both ranks (global chain order)
    assert open(out1 + ".pic.ave.dat").read() == open(out2 + ".pic.ave.dat").read()
    from mcdiff_b200.log import load_logger
    l1, l2 = load_logger(out1 + ".pic"), load_logger(out2 + ".pic")
    assert l1.nchains == l2.nchains == 7 and l2.nf == 7 * 3
    np.testing.assert_array_equal(l1.pack_rows(), l2.pack_rows())


@pytest.mark.parametrize("mode", ["pull", "radial", "blocked"])
def test_two_ranks_equal_one_rank_other_paths(engine, gold_ll, tmp_path, mode):
    """Same check for the other kernels of the path: --pull (general engine), --rad (radial loop) and a 120-bin
    profile (blocked engine)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from mcdiff_b200.transitions import write_Tmat_cube, write_Tmat_square
    from oracle import mcdiff_oracle as O
    tmp = str(tmp_path)
    files = []
    if mode == "radial":
        RR = np.load(os.path.join(ROOT, "tests", "golden", "radial_replay_golden.npz"))
        counts, lags = RR["radcos/counts"], RR["radcos/lags"]
        nr, dim_rad, lmax = counts.shape[-1], int(RR["radcos/dim_rad"]), int(RR["radcos/lmax"])
        for il, lag in enumerate(lags):
            fn = os.path.join(tmp, "rad.%d.dat" % il)
            write_Tmat_cube(counts[il].astype(int), fn, float(lag), "pbc", edges=np.linspace(-4.0, 4.0, nr + 1),
                            redges=np.arange(dim_rad) * 0.5, dt=1.0, dn=int(lag))
            files.append(fn)
        common = ["--model", "RadCosinusModel", "--ncosDrad", "4", "--rad", str(lmax), "-n", "100", "--nmc_update", "40",
                  "-T", "1", "--Tend", "1", "--dwrad", "0.4", "-s", "5", "--nchains", "5"]
        keys = ("log_like", "wrad")
    else:
        if mode == "blocked":
            lags = np.array([5.0, 10.0])
            counts, _, edges, _, _ = O.synthetic_counts(n=120, lags=lags, seed=2, total=1.0e5, dz=0.5)
            n = 120
        else:
            G = gold_ll
            counts, lags, edges, n = G["P_counts"], G["P_lags"], G["P_edges"], 100
        for l, lag in enumerate(lags):
            fn = os.path.join(tmp, "transitions.nbins%d.%d.pbc.dat" % (n, int(lag)))
            write_Tmat_square(counts[l], fn, float(lag), "pbc", edges=edges, dt=1.0, dn=int(lag))
            files.append(fn)
        common = ["--model", "CosinusModel", "--ncosF", "6", "--ncosD", "4", "-n", "100", "--nmc_update", "50", "-T", "1",
                  "--Tend", "1", "--dv", "0.1", "--dw", "0.1", "-s", "11", "--nchains", "5"]
        if mode == "pull":
            common += ["--pull", "0.3"]
        keys = ("log_like", "naccv", "naccw", "v_param", "w_param")
    script = os.path.join(ROOT, "scripts", "run-mcdiff")
    out1, out2 = os.path.join(tmp, "one.dat"), os.path.join(tmp, "two.dat")
    env = dict(os.environ)
    r1 = subprocess.run([sys.executable, script, "-o", out1] + common + files, capture_output=True, text=True, env=env)
    assert r1.returncode == 0, r1.stderr[-2000:]
    r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                         "--master-addr", "127.0.0.1", "--master-port", "29534", script, "-o", out2] + common + files,
                        capture_output=True, text=True, env=env)
    assert r2.returncode == 0, r2.stderr[-2000:]
    a, b = np.load(out1 + ".chains.npz"), np.load(out2 + ".chains.npz")
    for key in keys:
        np.testing.assert_array_equal(a[key], b[key])
    assert open(out1).read() == open(out2).read()
    assert open(out1 + ".pic.ave.dat").read() == open(out2 + ".pic.ave.dat").read()


def test_lag_scan_two_ranks_equal_one_rank(engine, gold_ll, tmp_path):
    """BASELINE config 4: the (problem, chain) pairs of a lag scan sharded over two ranks give the files of a
    one-GPU scan (per-lag .dat of chain 0, every chain's final state, the mcdiff-infty extrapolation)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from mcdiff_b200.transitions import write_Tmat_square
    G = gold_ll
    tmp = str(tmp_path)
    files = []
    for l, lag in enumerate(G["P_lags"][:3]):
        fn = os.path.join(tmp, "transitions.nbins100.%d.pbc.dat" % int(lag))
        write_Tmat_square(G["P_counts"][l], fn, float(lag), "pbc", edges=G["P_edges"], dt=1.0, dn=int(lag))
        files.append(fn)
    common = ["--ncosF", "6", "--ncosD", "4", "-n", "200", "--nmc_update", "100", "-T", "1", "--Tend", "1",
              "--dv", "0.2", "--dw", "0.2", "-s", "21", "--nchains", "5", "hexd"] + files      # 15 pairs: 8 + 7
    script = os.path.join(ROOT, "scripts", "mcdiff-lagscan")
    d1, d2 = os.path.join(tmp, "one"), os.path.join(tmp, "two")
    env = dict(os.environ)
    r1 = subprocess.run([sys.executable, script, "-o", d1] + common, capture_output=True, text=True, env=env)
    assert r1.returncode == 0, r1.stderr[-2000:]
    r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                         "--master-addr", "127.0.0.1", "--master-port", "29535", script, "-o", d2] + common,
                        capture_output=True, text=True, env=env)
    assert r2.returncode == 0, r2.stderr[-2000:]
    names = ["out.nbins100.%d.pbc.dat" % l for l in (10, 20, 30)] + ["hexd.Dreal.dat", "hexd.t0.dat"]
    for nm in names:
        assert open(os.path.join(d1, nm)).read() == open(os.path.join(d2, nm)).read(), nm
    for nm in names[:3]:
        a, b = np.load(os.path.join(d1, nm + ".chains.npz")), np.load(os.path.join(d2, nm + ".chains.npz"))
        assert a["log_like"].shape == (5,)
        for key in ("log_like", "v", "w", "naccv", "naccw", "status"):
            np.testing.assert_array_equal(a[key], b[key])
        assert len(np.unique(a["log_like"])) == 5

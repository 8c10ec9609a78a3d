"""End to end through the reference-facing host API (find_parameters / MCState) on the GPU:
output files in the reference's formats, restart through --initf, chain 0 equal to the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _files(tmp, G, key="P"):
    from mcdiff_b200.transitions import write_Tmat_square
    files = []
    for l, lag in enumerate(G[key + "_lags"]):
        fn = os.path.join(tmp, "transitions.nbins100.%d.pbc.dat" % int(lag))
        write_Tmat_square(G[key + "_counts"][l], fn, float(lag), "pbc", edges=G[key + "_edges"], dt=1.0, dn=int(lag))
        files.append(fn)
    return files


def test_find_parameters_outputs_and_restart(engine, gold_ll, tmp_path, capsys):
    from mcdiff_b200 import outreading as outr
    from mcdiff_b200 import philox
    from mcdiff_b200.log import load_logger
    from mcdiff_b200.mc import find_parameters
    from oracle import mcdiff_oracle as O
    G = gold_ll
    tmp = str(tmp_path)
    files = _files(tmp, G)
    out = os.path.join(tmp, "out.nbins100.pbc.dat")
    nmc, nch, seed = 300, 8, 3
    MC, logger = find_parameters(files, True, "CosinusModel", 0.05, 0.05, 0.5, 1.0, 0.1, 1.0, 1.0, nmc, 100.0, seed,
                                 out, 10, 6, 0, False, None, -1, -1, False, None, nchains=nch)
    text = capsys.readouterr().out
    ll_first = float([l for l in text.splitlines() if l.startswith("initial log-likelihood:")][0].split()[-1])
    assert ll_first == pytest.approx(float(G["P_flat_ll"]), rel=1e-9)
    for suffix in ("", ".pic", ".pic.ave.dat", ".chains.npz"):
        assert os.path.exists(out + suffix), suffix

    # chain 0 == oracle replay on the tape the device Philox stream produces
    n = 100
    vb, wb = O.cos_basis_center(n, 10), O.cos_basis_border(n, n, 6)
    wc0 = np.zeros(6)
    wc0[0] = G["P_flat_w"][0]
    cfg = O.ChainConfig(G["P_counts"], G["P_lags"], True, v_basis=vb, w_basis=wb, dv=0.05, dw=0.05, temp=1.0,
                        temp_end=1.0, nmc=nmc, num_mc_update=100.0)
    tu, ti = philox.chain_tape(seed, 0, 0, nmc, n, n, 10, 6)
    ref = O.replay(cfg, np.zeros(n), wb @ wc0, tu, ti, v_coeff0=np.zeros(10), w_coeff0=wc0)
    assert (MC.naccv, MC.naccw) == (ref["naccv"], ref["naccw"])
    np.testing.assert_allclose(MC.log_like, ref["ll"], rtol=1e-9)
    np.testing.assert_allclose(MC.model.v_coeff, ref["v_coeff"], atol=1e-13)
    np.testing.assert_allclose([MC.dv, MC.dw], [ref["dv"], ref["dw"]], rtol=1e-13)

    # `.dat` = chain 0 in the reference's text format
    np.testing.assert_allclose(outr.read_Fcoeffs(out, final=True), ref["v_coeff"], rtol=1e-8, atol=1e-12)
    F, D, edges = outr.read_F_D_edges(out)
    np.testing.assert_allclose(F, ref["v"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(D, np.exp(ref["w"] + MC.model.wunit), rtol=1e-8)
    dv, dw = outr.read_dv_dw(out, final=True)
    assert dv == pytest.approx(ref["dv"]) and dw == pytest.approx(ref["dw"])

    # `.pic` = pooled Logger: nch * (nmc/100 + 1) rows, chain 0 rows equal the oracle's samples
    lg = load_logger(out + ".pic")
    assert lg.nf == nch * 4 and lg.v_coeff.shape == (nch * 4, 10)
    np.testing.assert_allclose(lg.log_like[:4], ref["samples"][:, 0], rtol=1e-9)
    np.testing.assert_allclose(lg.v_coeff[:4], ref["samples"][:, 6:16], atol=1e-13)
    Fa, Da, ea = outr.read_F_D_edges(out + ".pic.ave.dat")
    np.testing.assert_allclose(Fa, (lg.v_coeff @ vb.T).mean(axis=0), rtol=1e-7, atol=1e-9)
    ch = np.load(out + ".chains.npz")
    assert ch["log_like"].shape == (nch,) and not ch["status"].any()
    assert len(np.unique(ch["log_like"])) == nch              # independent streams
    np.testing.assert_allclose(ch["log_like"][0], ref["ll"], rtol=1e-9)

    # restart: --initf reads the final coefficients back (lib/MCState.py:139-201)
    out2 = os.path.join(tmp, "run2.dat")
    MC2, _ = find_parameters(files, True, "CosinusModel", 0.05, 0.05, 0.5, 1.0, 0.1, 1.0, 1.0, 100, 100.0, 5,
                             out2, 10, 6, 0, False, out, -1, -1, False, None, nchains=2)
    text2 = capsys.readouterr().out
    ll_init = float([l for l in text2.splitlines() if l.startswith("initial log-likelihood:")][0].split()[-1])
    assert ll_init == pytest.approx(ref["ll"], rel=2e-7)      # coefficients are stored with 9 digits


def test_exact_resume_from_chainfile(engine, gold_ll, tmp_path, capsys):
    """<outfile>.chains.npz holds the whole per-chain state and the Philox counters, so
    200 steps == 100 steps + `--resume` + 100 steps, bit for bit, for every chain."""
    from mcdiff_b200.mc import find_parameters
    G = gold_ll
    tmp = str(tmp_path)
    files = _files(tmp, G)
    args = (files, True, "CosinusModel", 0.05, 0.05, 0.5, 1.0, 0.1, 1.0, 1.0)
    tail = (50.0, 9)
    kw = dict(nchains=5)
    full, a, b = (os.path.join(tmp, x) for x in ("full.dat", "a.dat", "b.dat"))
    find_parameters(*args, 200, *tail, full, 10, 6, 0, False, None, -1, -1, False, None, **kw)
    find_parameters(*args, 100, *tail, a, 10, 6, 0, False, None, -1, -1, False, None, **kw)
    find_parameters(*args, 100, 50.0, 1234, b, 10, 6, 0, False, None, -1, -1, False, None, resume=a + ".chains.npz", **kw)
    capsys.readouterr()
    zf, zb = np.load(full + ".chains.npz"), np.load(b + ".chains.npz")
    for key in ("scal", "counters", "v", "w", "v_coeff", "w_coeff", "naccv_coeff", "naccw_coeff"):
        np.testing.assert_array_equal(zf[key], zb[key], err_msg=key)
    assert int(zb["steps_done"]) == 200 and int(zb["seed"]) == 9
    # the final file is the same except for the line that echoes this invocation's --nmc (100 vs 200)
    strip = lambda fn: [l for l in open(fn).read().splitlines() if not l.startswith("n(MC)=")]  # noqa: E731
    assert strip(full) == strip(b)


def test_per_bin_model_and_nopbc(engine, gold_ll, tmp_path, capsys):
    from mcdiff_b200 import outreading as outr
    from mcdiff_b200.mc import find_parameters
    from mcdiff_b200.transitions import write_Tmat_square
    G = gold_ll
    tmp = str(tmp_path)
    fn = os.path.join(tmp, "cut.dat")
    write_Tmat_square(G["C_counts"][0], fn, 20.0, "cut", edges=G["C_edges"], dt=1.0, dn=20)
    out = os.path.join(tmp, "cut.out.dat")
    MC, lg = find_parameters([fn], False, "Model", 0.1, 0.1, 0.5, 1.0, 0.1, 1.0, 1.0, 200, 100.0, 1, out,
                             0, 0, 0, False, None, -1, -1, False, None, nchains=4)
    capsys.readouterr()
    F, D, edges = outr.read_F_D_edges(out)
    assert len(F) == 60 and len(D) == 59
    assert lg.v.shape == (4 * 3, 60) and lg.w.shape == (4 * 3, 59)
    assert MC.log_like > float(MC.chain_log_like.min()) - 1e9 and not MC.chain_status.any()
    with pytest.raises(AssertionError):      # pbc flag must agree with the counting method (lib/MCState.py:99)
        find_parameters([fn], True, "Model", 0.1, 0.1, 0.5, 1.0, 0.1, 1.0, 1.0, 10, 100.0, 1, None,
                        0, 0, 0, False, None, -1, -1, False, None)
    with pytest.raises(NotImplementedError):
        find_parameters([fn], False, "Model", 0.1, 0.1, 0.5, 1.0, 0.1, 1.0, 1.0, 10, 100.0, 1, None,
                        0, 0, 0, False, None, -1, -1, False, 0.1)


def test_initial_nan_raises_like_reference(engine, gold_ll):
    """lib/MCState.py:114-117: a non-finite first likelihood is a ValueError."""
    from mcdiff_b200.engine import ChainBatch, DeviceProblem
    G = gold_ll
    prob = DeviceProblem(engine, G["P_counts"], G["P_lags"], True)
    ch = ChainBatch(prob, 2)
    v = np.zeros((2, 100))
    v[1, 0] = np.inf
    ch.set_state(v, np.tile(G["P_flat_w"], (2, 1)), dv=0.1, dw=0.1)
    with pytest.raises(ValueError, match="Initial likelihood diverges"):
        ch.init_log_like()


def test_radial_find_parameters(engine, tmp_path, capsys):
    """run-mcdiff --rad LMAX --model RadCosinusModel: radial MC on the device, `===== final Drad =====`
    section readable by read_Drad, chain 0 equal to the reference trajectory fixture's start."""
    from mcdiff_b200 import outreading as outr
    from mcdiff_b200.log import load_logger
    from mcdiff_b200.mc import find_parameters
    from mcdiff_b200.transitions import write_Tmat_cube
    RR = np.load(os.path.join(os.path.dirname(__file__), "golden", "radial_replay_golden.npz"))
    g = lambda k: RR["radcos/" + k]  # noqa: E731
    counts, lags = g("counts"), g("lags")
    nr, dim_rad, lmax = counts.shape[-1], int(g("dim_rad")), int(g("lmax"))
    tmp = str(tmp_path)
    edges = np.linspace(-4.0, 4.0, nr + 1)
    redges = np.arange(dim_rad) * 0.5
    files = []
    for il, lag in enumerate(lags):
        fn = os.path.join(tmp, "rad.%d.dat" % il)
        write_Tmat_cube(counts[il].astype(int), fn, float(lag), "pbc", edges=edges, redges=redges, dt=1.0, dn=int(lag))
        files.append(fn)
    out = os.path.join(tmp, "rad.out.dat")
    MC, lg = find_parameters(files, True, "RadCosinusModel", 0.5, 0.5, 0.4, 1.0, 0.1, 1.0, 1.0, 200, 40.0, 9, out,
                             0, 0, 4, False, None, -1, lmax, False, None, nchains=3)
    text = capsys.readouterr().out
    ll_first = float([l for l in text.splitlines() if l.startswith("initial log-likelihood:")][0].split()[-1])
    assert ll_first == pytest.approx(float(g("ll0")), rel=1e-9)        # same start as the reference fixture
    assert MC.log_like > ll_first and MC.naccwrad > 0 and not MC.chain_status.any()
    Drad, redg = outr.read_Drad(out)
    np.testing.assert_allclose(Drad, np.exp(MC.model.wrad + MC.model.wradunit), rtol=1e-8)
    assert "===== final wrad_coeff =====" in open(out).read()
    lgr = load_logger(out + ".pic")
    assert lgr.wrad_coeff.shape == (3 * 3, 4)
    Da, _ = outr.read_Drad(out + ".pic.ave.dat")
    assert len(Da) == nr
    ch = np.load(out + ".chains.npz")
    assert ch["wrad"].shape == (3, nr) and len(np.unique(ch["log_like"])) == 3


def test_lag_scan_concurrent_streams(engine, gold_ll, tmp_path, capsys):
    """BASELINE config 4 shape: independent single-lag problems on concurrent CUDA streams, then the
    mcdiff-infty extrapolation.  Each problem's chain 0 equals the same problem run alone."""
    from mcdiff_b200.infty import extrapolate, run_lag_scan
    from mcdiff_b200.mc import find_parameters
    from mcdiff_b200 import outreading as outr
    G = gold_ll
    tmp = str(tmp_path)
    files = _files(tmp, G)
    by_lag = {float(l): f for l, f in zip(G["P_lags"], files)}
    outs = run_lag_scan(by_lag, os.path.join(tmp, "scan"), ncosF=6, ncosD=4, dv=0.2, dw=0.2, nmc=200, nmc_update=100.0,
                        seed=21, nchains=64)
    assert [os.path.basename(o) for o in outs] == ["out.nbins100.%d.pbc.dat" % l for l in (10, 20, 30, 40)]
    # problem k alone (seed + k), same chain 0
    alone = os.path.join(tmp, "alone.dat")
    find_parameters([files[2]], True, "CosinusModel", 0.2, 0.2, 0.5, 1.0, 0.1, 1.0, 1.0, 200, 100.0, 23, alone,
                    6, 4, 0, False, None, -1, -1, False, None, nchains=64)
    capsys.readouterr()
    np.testing.assert_array_equal(outr.read_Fcoeffs(outs[2], final=True), outr.read_Fcoeffs(alone, final=True))
    np.testing.assert_array_equal(outr.read_F_D_edges(outs[2])[1], outr.read_F_D_edges(alone)[1])
    lts, Dreal, t0 = extrapolate(outs, os.path.join(tmp, "hexd."))
    np.testing.assert_array_equal(lts, [10.0, 20.0, 30.0, 40.0])
    assert Dreal.shape == (100,) and np.all(np.isfinite(Dreal))


def test_find_parameters_with_pull(engine, gold_pull, tmp_path, capsys):
    """run-mcdiff --pull through the host API: dF = -pull dz as lib/MCState.py:65-67, first likelihood equal to
    the reference's, chain 0 equal to the oracle replay of the device Philox tape."""
    from mcdiff_b200 import philox
    from mcdiff_b200.mc import find_parameters
    from mcdiff_b200.transitions import write_Tmat_square
    from oracle import mcdiff_oracle as O
    G = gold_pull
    tmp = str(tmp_path)
    files = []
    for l, lag in enumerate(G["lags"]):
        fn = os.path.join(tmp, "transitions.nbins100.%d.pbc.dat" % int(lag))
        write_Tmat_square(G["counts"][l], fn, float(lag), "pbc", edges=G["edges"], dt=1.0, dn=int(lag))
        files.append(fn)
    out = os.path.join(tmp, "out.pull.dat")
    nmc, seed, pull = 120, 5, float(G["cosPull/cfg_pull"])
    MC, logger = find_parameters(files, True, "CosinusModel", 0.3, 0.3, 0.5, 1.0, 0.5, 1.0, 1.0, nmc, 100.0, seed, out,
                                 6, 4, 0, True, None, -1, -1, False, pull, nchains=4)
    text = capsys.readouterr().out
    assert MC.pull == pytest.approx(float(G["cosPull/dF"]), rel=1e-12)
    ll_first = float([l for l in text.splitlines() if l.startswith("initial log-likelihood:")][0].split()[-1])
    assert ll_first == pytest.approx(float(G["cosPull/ll0"]), rel=1e-9)
    n = 100
    vb, wb = O.cos_basis_center(n, 6), O.cos_basis_border(n, n, 4)
    cfg = O.ChainConfig(G["counts"], G["lags"], True, v_basis=vb, w_basis=wb, dv=0.3, dw=0.3, dtimezero=0.5, temp=1.0,
                        temp_end=1.0, nmc=nmc, num_mc_update=100.0, move_timezero=True, pull=MC.pull)
    tu, ti = philox.chain_tape(seed, 0, 0, nmc, n, n, 6, 4)
    ref = O.replay(cfg, G["cosPull/v0"], G["cosPull/w0"], tu, ti, v_coeff0=G["cosPull/vc0"], w_coeff0=G["cosPull/wc0"])
    assert (MC.naccv, MC.naccw, MC.nacctimezero) == (ref["naccv"], ref["naccw"], ref["nacct0"])
    np.testing.assert_allclose(MC.log_like, ref["ll"], rtol=1e-9)
    assert os.path.exists(out)


@pytest.mark.parametrize("model,ncosF,ncosD", [("SinusCosinusModel", 5, 3), ("StepModel", 4, 2), ("OneStepModel", 4, 2)])
def test_other_model_families_on_device(engine, gold_ll, tmp_path, capsys, model, ncosF, ncosD):
    """The kernel takes the basis matrices of any model (lib/model.py:169-185, 238-262, 303-310): chain 0 equals the
    oracle replay with the same bases on the device Philox tape."""
    from mcdiff_b200 import philox
    from mcdiff_b200.mc import find_parameters
    from oracle import mcdiff_oracle as O
    G = gold_ll
    files = _files(str(tmp_path), G)
    out = os.path.join(str(tmp_path), "out.%s.dat" % model)
    nmc, seed = 120, 9
    MC, logger = find_parameters(files, True, model, 0.2, 0.2, 0.5, 1.0, 0.1, 1.0, 1.0, nmc, 100.0, seed, out,
                                 ncosF, ncosD, 0, False, None, -1, -1, False, None, nchains=3)
    capsys.readouterr()
    m = MC.model
    n = 100
    from mcdiff_b200.model import MODELS_1D
    from mcdiff_b200.transitions import Transitions
    m0 = MODELS_1D[model](Transitions(files), 1.0, ncosF, ncosD, 0)        # the start state the model class defines
    vc0, wc0 = m0.v_coeff.copy(), m0.w_coeff.copy()
    cfg = O.ChainConfig(G["P_counts"], G["P_lags"], True, v_basis=m.v_basis, w_basis=m.w_basis, dv=0.2, dw=0.2, temp=1.0,
                        temp_end=1.0, nmc=nmc, num_mc_update=100.0)
    tu, ti = philox.chain_tape(seed, 0, 0, nmc, n, n, ncosF, ncosD)
    ref = O.replay(cfg, m.v_basis @ vc0, m.w_basis @ wc0, tu, ti, v_coeff0=vc0, w_coeff0=wc0)
    assert (MC.naccv, MC.naccw) == (ref["naccv"], ref["naccw"])
    np.testing.assert_allclose(MC.log_like, ref["ll"], rtol=1e-9)
    np.testing.assert_allclose(MC.model.v_coeff, ref["v_coeff"], atol=1e-13)


def test_posterior_statistics_on_the_device_equal_numpy(engine):
    """Many chains: the Logger forms profile rows, exponentials and their mean / std on the GPU
    (Logger._device_stats); same numbers as the numpy formulas of lib/log.py:90-106 to rounding."""
    from mcdiff_b200.log import Logger
    rng = np.random.default_rng(1)
    rows, ncos, n = 20000, 6, 100
    coeff = rng.normal(0, 0.3, size=(rows, ncos))
    basis = rng.normal(0, 1.0, size=(n, ncos))
    lg = Logger()
    d = lg._device_stats(coeff, basis, True, True)
    assert d is not None
    a = coeff @ basis.T
    ea = np.exp(a)
    for got, want in zip(d, (np.mean(a, 0), np.std(a, 0), np.mean(ea, 0), np.std(ea, 0))):
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)
    assert lg._device_stats(coeff[:100], basis, True, True) is None          # few rows stay on the host

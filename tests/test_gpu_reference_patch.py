"""INTEGRATION.md 1 applied to the UNMODIFIED reference: integration/b200.py (the file a maintainer adds) replaces
`log_like_lag` inside the real `mcdiff.MCState` module (installed under baseline/_ref by
scripts/install_reference.sh; it travels to the GPU box, /root/reference does not).  The reference's own
init_log_like / mcmove_* then run on the GPU likelihood and must reproduce the archived values K1-K4 and the
decisions of its scipy path."""
import importlib.util
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.fixture(scope="module")
def ref_mcdiff(engine):
    if not os.path.isdir(os.path.join(REF, "mcdiff")):
        pytest.skip("baseline/_ref not installed (scripts/install_reference.sh)")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    os.environ["MCDIFF_B200_LIB"] = os.path.join(ROOT, "mcdiff_b200", "libmcdiff_b200.so")
    spec = importlib.util.spec_from_file_location("mcdiff.b200", os.path.join(ROOT, "integration", "b200.py"))
    b200 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b200)
    import mcdiff.MCState as RM
    import mcdiff.tools.extract as RX
    import mcdiff.transitions as RT
    return b200, RM, RT, RX


def _ref_state(RM, RT, RX, tmp, counts, lags, edges, ncosF, ncosD, seed=None, move_timezero=False):
    files = []
    for l, lag in enumerate(lags):
        fn = os.path.join(tmp, "transitions.%d.dat" % l)
        RX.write_Tmat_square(counts[l], fn, float(lag), "pbc", edges=edges, dt=1.0, dn=int(lag))
        files.append(fn)
    MC = RM.MCState(True, -1)
    MC.set_MC_params(0.05, 0.05, 0.5, 1.0, 0.1, 1.0, 100, 50.0, move_timezero, -1, temp_end=1.0)
    MC.set_model("CosinusModel", RT.Transitions(files), ncosF, ncosD, 0, None)
    return MC


def test_patched_reference_reproduces_archived_values(ref_mcdiff, gold_ll, tmp_path, capsys):
    b200, RM, RT, RX = ref_mcdiff
    G = gold_ll
    # K1, K2 were logged on transition matrix A, K3, K4 on M (tests/golden/make_golden.py)
    for fixture, ks in (("A", (0, 1)), ("M", (2, 3))):
        tmp = os.path.join(str(tmp_path), fixture)
        os.makedirs(tmp)
        MC = _ref_state(RM, RT, RX, tmp, G[fixture + "_counts"], G[fixture + "_lags"], G["A_edges"], 0, 0)
        orig = RM.log_like_lag
        RM.log_like_lag = b200.B200Likelihood(MC.data, MC.pbc, pull=MC.pull)     # the patch of INTEGRATION.md 1
        try:
            for k in ks:
                MC.model.v = G["K_v"][k].copy()
                MC.model.w = G["K_w"][k].copy()
                MC.init_log_like()
                assert MC.log_like == pytest.approx(float(G["K_archived"][k]), rel=1e-9)    # as printed in the logs
                assert MC.log_like == pytest.approx(float(G["K_ll"][k]), rel=1e-11)         # the reference's scipy path
        finally:
            RM.log_like_lag.close()
            RM.log_like_lag = orig
    capsys.readouterr()


def test_patched_reference_moves_equal_scipy_path(ref_mcdiff, gold_ll, tmp_path, capsys):
    """The reference's own mcmove_potential / mcmove_diffusion / mcmove_timezero on the GPU likelihood make the
    decisions of its scipy path (same np.random seed), including the shifted lag times of --t0 trials."""
    b200, RM, RT, RX = ref_mcdiff
    G = gold_ll

    def trajectory(patched):
        MC = _ref_state(RM, RT, RX, str(tmp_path), G["P_counts"], G["P_lags"], G["P_edges"], 6, 4, move_timezero=True)
        orig = RM.log_like_lag
        if patched:
            RM.log_like_lag = b200.B200Likelihood(MC.data, MC.pbc, pull=MC.pull)
        try:
            np.random.seed(11)
            MC.init_log_like()
            lls = [MC.log_like]
            for imc in range(40):
                if np.random.rand() < 0.5:
                    MC.mcmove_potential()
                else:
                    MC.mcmove_diffusion()
                MC.mcmove_timezero()
                lls.append(MC.log_like)
            return np.array(lls), (MC.naccv, MC.naccw, MC.nacctimezero), MC.model.timezero
        finally:
            if patched:
                RM.log_like_lag.close()
            RM.log_like_lag = orig

    ll_g, acc_g, t0_g = trajectory(True)
    ll_c, acc_c, t0_c = trajectory(False)
    capsys.readouterr()
    assert acc_g == acc_c and t0_g == t0_c and acc_g[2] > 0
    np.testing.assert_allclose(ll_g, ll_c, rtol=1e-11)

"""Host-side mirror of the reference interface: input format, models, output format."""
import io
import os

import numpy as np
import pytest

from mcdiff_b200 import log as mlog
from mcdiff_b200 import outreading as outr
from mcdiff_b200.model import CosinusModel, Model, OneStepModel, SinusCosinusModel, StepModel
from mcdiff_b200.parallel import shard_range
from mcdiff_b200.transitions import (RadTransitions, Transitions, reduce_Tmat, write_Tmat_cube,
                                     write_Tmat_square)
from oracle import mcdiff_oracle as O


def _write_files(tmp_path, gold_ll, key="P"):
    G = gold_ll
    files = []
    for l, lag in enumerate(G[key + "_lags"]):
        fn = os.path.join(tmp_path, "transitions.nbins100.%d.pbc.dat" % int(lag))
        write_Tmat_square(G[key + "_counts"][l], fn, float(lag), "pbc", edges=G[key + "_edges"], dt=1.0, dn=int(lag))
        files.append(fn)
    return files


def test_transitions_roundtrip(tmp_path, gold_ll):
    files = _write_files(str(tmp_path), gold_ll)
    tr = Transitions(files)
    assert tr.dim_lt == 4 and tr.dim_trans == 100 and tr.count == "pbc"
    assert np.array_equal(tr.list_trans, gold_ll["P_counts"])
    np.testing.assert_allclose(tr.list_lt, [10, 20, 30, 40])
    assert tr.min_lt == 10.0
    np.testing.assert_allclose(tr.edges, gold_ll["P_edges"])
    assert tr.dz == pytest.approx((tr.edges[-1] - tr.edges[0]) / 100.0)
    # the oracle's minimal reader parses the same file
    h, m = O.read_counts(files[1])
    assert h["lt"] == 20.0 and np.array_equal(m, gold_ll["P_counts"][1])


def test_header_rules(tmp_path):
    fn = os.path.join(str(tmp_path), "a.dat")
    with open(fn, "w") as f:
        f.write("#dt 2.0\n#dn 5\n#count cut\n1 2\n3 4\n")
    tr = Transitions([fn])
    assert tr.list_lt[0] == 10.0 and tr.count == "cut"
    assert np.array_equal(tr.edges, [0.0, 1.0, 2.0]) and tr.dz == 1.0
    with open(fn, "w") as f:
        f.write("#lt 3.0\n#dt 1.0\n#dn 5\n#count pbc\n1 2\n3 4\n")
    with pytest.raises(AssertionError):
        Transitions([fn])
    with open(fn, "w") as f:
        f.write("#lt 3.0\n#dt 1.0\n1 2\n3 4\n")
    with pytest.raises(ValueError):
        Transitions([fn])                     # no #count
    with open(fn, "w") as f:
        f.write("#lt 3.0\n#dt 1.0\n#count pbc\n1 2 3\n3 4 5\n")
    with pytest.raises(ValueError):
        Transitions([fn])                     # ragged: 2 rows of 3


def test_reduction_marks_cut(gold_ll):
    A = np.zeros((6, 6), dtype=np.int64)
    A[1:5, 1:5] = gold_ll["A_counts"][0][:4, :4] + 1
    header = dict(count="pbc", edges=np.arange(7.0), lt=1.0)
    dim, h2, T = reduce_Tmat(6, header, A)
    assert dim == 4 and h2["count"] == "cut" and T.shape == (4, 4)
    np.testing.assert_allclose(h2["edges"], [1, 2, 3, 4, 5])
    assert header["count"] == "pbc"           # caller's header untouched


def test_rad_transitions_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    cube = rng.integers(0, 50, size=(3, 5, 5))
    fn = os.path.join(str(tmp_path), "cube.dat")
    write_Tmat_cube(cube, fn, 4.0, "pbc", edges=np.arange(6.0), redges=np.arange(0, 1.5, 0.5), dt=1.0, dn=4)
    tr = RadTransitions([fn])
    assert tr.dim_rad == 3 and tr.dim_trans == 5
    assert np.array_equal(tr.list_trans[0], cube)
    np.testing.assert_allclose(tr.redges, [0.0, 0.5, 1.0])


def test_models_match_oracle_bases(gold_ll):
    tr = Transitions.from_arrays(gold_ll["P_counts"], gold_ll["P_lags"], gold_ll["P_edges"], "pbc")
    m = CosinusModel(tr, 1.0, 10, 6)
    np.testing.assert_allclose(m.v_basis, O.cos_basis_center(100, 10), rtol=0, atol=0)
    np.testing.assert_allclose(m.w_basis, O.cos_basis_border(100, 100, 6), rtol=0, atol=0)
    dz = tr.edges[1] - tr.edges[0]
    assert m.wunit == pytest.approx(np.log(dz * dz))
    np.testing.assert_allclose(m.w, -m.wunit)                  # D0 = 1 A^2/ps
    np.testing.assert_allclose(m.w, gold_ll["P_flat_w"], rtol=1e-14)
    assert m.pbc and m.dim_w == 100
    s = SinusCosinusModel(tr, 2.0, 5, 3)
    np.testing.assert_allclose(s.v_basis, O.sincos_basis_center(100, 5))
    np.testing.assert_allclose(s.w_basis, O.sincos_basis_border(100, 100, 3))
    with pytest.raises(AssertionError):
        SinusCosinusModel(tr, 1.0, 4, 3)
    st = StepModel(tr, 1.0, 4, 2)
    assert st.v_basis.shape == (100, 4) and set(np.unique(st.v_basis)) == {0.0, 1.0}
    assert np.all(st.v_basis[:, 0] == 1.0) and st.v_basis[:, 1].sum() == 76
    one = OneStepModel(tr, 1.0, 4, 2)
    assert one.v_basis[:, 1].sum() == 75 and one.w_coeff[1] == 0.0
    cut = Transitions.from_arrays(gold_ll["C_counts"], gold_ll["C_lags"], gold_ll["C_edges"], "cut")
    mc = Model(cut, 1.0)
    assert not mc.pbc and mc.dim_w == mc.dim_v - 1


class _MC:
    do_radial = False
    nmc = 200
    dv, dw, dwrad, temp, num_MC_update, k = 0.05, 0.04, 0.5, 1.0, 100.0, -1


def test_result_file_roundtrip(tmp_path, gold_ll):
    """print_coeffs / print_profiles output is parsed back by the outreading functions
    (the --initf restart path of lib/MCState.py:139-201)."""
    tr = Transitions.from_arrays(gold_ll["P_counts"], gold_ll["P_lags"], gold_ll["P_edges"], "pbc")
    m = CosinusModel(tr, 1.0, 5, 3)
    rng = np.random.default_rng(1)
    m.v_coeff = rng.normal(size=5)
    m.w_coeff[1:] = rng.normal(size=2) * 0.1
    m.update_v()
    m.update_w()
    fn = os.path.join(str(tmp_path), "out.dat")
    with open(fn, "w") as f:
        print("----- final Settings MC -----", file=f)
        print("dv(MC-potential)=", 0.0421, file=f)
        print("dw(MC-logD)=", 0.0377, file=f)
        print("-" * 20, file=f)
        mlog.print_coeffs(f, m, m.v_coeff, m.w_coeff, None, None, final=True)
        mlog.print_profiles(f, m, m.v, m.w, final=True)
    text = open(fn).read()
    assert "===== final v_coeff =====" in text and "===== final F D =====" in text and text.endswith("#Done\n")
    assert "   index  bin-str  bin-end" in text
    np.testing.assert_allclose(outr.read_Fcoeffs(fn, final=True), m.v_coeff, rtol=1e-8)
    wc = outr.read_Dcoeffs(fn, final=True)
    assert wc[0] == pytest.approx(m.w_coeff[0] + m.wunit, rel=1e-8)       # flat term stored shifted
    F, D, edges = outr.read_F_D_edges(fn)
    np.testing.assert_allclose(F, m.v, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(D, np.exp(m.w + m.wunit), rtol=1e-8)
    np.testing.assert_allclose(edges, np.round(tr.edges, 3), atol=1e-9)
    assert outr.read_dv_dw(fn, final=True) == (0.0421, 0.0377)
    row = text.split("bin-end")[1].splitlines()[1]
    assert row == "%8d %8.3f %8.3f  %16.8e %16.8e" % (0, tr.edges[0], tr.edges[1], m.v[0], np.exp(m.w[0] + m.wunit))


def test_nopbc_profile_file(tmp_path, gold_ll):
    cut = Transitions.from_arrays(gold_ll["C_counts"], gold_ll["C_lags"], gold_ll["C_edges"], "cut")
    m = Model(cut, 1.0)
    m.v = np.linspace(0, 1, m.dim_v)
    fn = os.path.join(str(tmp_path), "cut.dat")
    with open(fn, "w") as f:
        mlog.print_profiles(f, m, m.v, m.w, final=True)
    F, D, edges = outr.read_F_D_edges(fn)
    assert len(F) == m.dim_v and len(D) == m.dim_v - 1         # last bin has F only


def test_logger_pooling_and_average(tmp_path, gold_ll):
    tr = Transitions.from_arrays(gold_ll["P_counts"], gold_ll["P_lags"], gold_ll["P_edges"], "pbc")
    m = CosinusModel(tr, 1.0, 4, 2)
    mc = _MC()
    mc.model = m
    lg = mlog.Logger(mc, nchains=3)
    assert lg.nf_chain == 3 and lg.nf == 9
    rng = np.random.default_rng(2)
    rows = rng.normal(size=(3, 3, 6 + 4 + 2))
    for c in range(3):
        lg.fill_rows(c, 0, rows[c, :1], m)
        lg.fill_rows(c, 1, rows[c, 1:], m)
    np.testing.assert_allclose(lg.v_coeff[lg.chain_rows(1)], rows[1, :, 6:10])
    np.testing.assert_allclose(lg.log_like.reshape(3, 3), rows[:, :, 0])
    lg.model = m
    pic = os.path.join(str(tmp_path), "o.pic")
    lg.dump(pic)
    lg2 = mlog.load_logger(pic)
    np.testing.assert_allclose(lg2.w_coeff, lg.w_coeff)
    mlog.write_average_from_pic(pic, pic + ".ave.dat")
    F, D, edges = outr.read_F_D_edges(pic + ".ave.dat")
    prof = lg.v_coeff @ m.v_basis.T
    np.testing.assert_allclose(F, prof.mean(axis=0), rtol=1e-7, atol=1e-9)
    wprof = lg.w_coeff @ m.w_basis.T
    np.testing.assert_allclose(D, np.exp(wprof).mean(axis=0) * np.exp(m.wunit), rtol=1e-7)   # mean of exp
    Fl, Dl, e, Fst, Dst = outr.read_F_D_edges_logger(lg2)
    np.testing.assert_allclose(Fst, prof.std(axis=0), rtol=1e-12)


def test_shard_range_covers_all_chains():
    for n in (1, 7, 1024, 4099):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_cli_keeps_reference_options():
    from mcdiff_b200.cli import build_parser
    p = build_parser()
    o, a = p.parse_args(["f.dat"])
    assert (o.nmc, o.temp, o.temp_end, o.dv, o.dw, o.dwrad, o.model) == (1000, 1.0, 0.01, 0.5, 0.5, 0.5, "Model")
    assert o.nmc_update == 100 and isinstance(p.parse_args(["--nmc_update", "50", "f"])[0].nmc_update, float)
    assert o.pbc is True and p.parse_args(["--nopbc", "f"])[0].pbc is False
    o2, a2 = p.parse_args("-o out -n 5000 -T 1 --Tend 1 --ncosF 10 --ncosD 6 --model CosinusModel -s 3 --t0 "
                          "--dt0 0.2 -k 2.5 --rad 20 --reduction --pull 0.1 --initf x.dat a b".split())
    assert (o2.outfile, o2.nmc, o2.ncosF, o2.ncosD, o2.seed, o2.move_timezero, o2.k, o2.lmax) == \
        ("out", 5000, 10, 6, 3, True, 2.5, 20)
    assert o2.reduction and o2.pull == 0.1 and o2.initfile == "x.dat" and a2 == ["a", "b"]
    assert p.parse_args(["--nchains", "4096", "f"])[0].nchains == 4096


def test_philox_tape_index_ranges():
    from mcdiff_b200 import philox
    u, idx = philox.chain_tape(1, 5, 0, 500, 100, 100, 10, 6)
    pot = u[:, 0] < 0.5
    assert idx[pot].min() >= 1 and idx[pot].max() <= 9          # flat basis function never moved
    assert idx[~pot].min() >= 0 and idx[~pot].max() <= 5
    assert np.all((u >= 0) & (u < 1))
    u2, idx2 = philox.chain_tape(1, 5, 100, 10, 100, 100, 10, 6)
    np.testing.assert_array_equal(u2, u[100:110])               # counter-based: resumable anywhere
    u3, _ = philox.chain_tape(1, 6, 0, 10, 100, 100, 10, 6)
    assert not np.array_equal(u3, u[:10])


def test_infty_fit_equals_reference_lstsq(tmp_path):
    """Closed-form per-bin fit == the reference's big lstsq (scripts/mcdiff-infty:99-117)."""
    from mcdiff_b200.infty import extract_lagtimes_from_filenames, fit_line_inverse, extrapolate
    rng = np.random.default_rng(4)
    lts = np.array([1.0, 2.0, 5.0, 10.0, 20.0, 50.0])
    nb = 12
    Dtrue, t0true = rng.uniform(0.2, 1.0, nb), rng.uniform(0.5, 3.0, nb)
    D = Dtrue[:, None] * (1.0 + t0true[:, None] / lts[None, :]) + rng.normal(0, 1e-3, (nb, len(lts)))
    A = np.zeros((nb * len(lts), 2 * nb))
    res = np.zeros(nb * len(lts))
    for i, lt in enumerate(lts):            # the reference's design matrix
        for j in range(nb):
            res[i * nb + j] = D[j, i]
            A[i * nb + j, j] = 1.0
            A[i * nb + j, j + nb] = 1.0 / lt
    x = np.linalg.lstsq(A, res, rcond=-1)[0]
    Dreal, t0 = fit_line_inverse(lts, D)
    np.testing.assert_allclose(Dreal, x[:nb], rtol=1e-10)
    np.testing.assert_allclose(t0, x[nb:] / x[:nb], rtol=1e-9)
    np.testing.assert_allclose(Dreal, Dtrue, atol=5e-3)
    names = ["/x/out.nbins100.%d.pbc.dat" % l for l in (1, 20, 50)]
    np.testing.assert_array_equal(extract_lagtimes_from_filenames(names), [1.0, 20.0, 50.0])
    # end to end on result files written in the reference's profile format
    from mcdiff_b200 import log as mlog

    class M:
        pbc, dim_v, wunit = True, nb, 0.0
        edges = np.arange(nb + 1.0)
    files = []
    for i, lt in enumerate(lts):
        fn = os.path.join(str(tmp_path), "out.nbins%d.%d.pbc.dat" % (nb, int(lt)))
        with open(fn, "w") as f:
            mlog.print_profiles(f, M, np.zeros(nb), np.log(D[:, i]), final=True)
        files.append(fn)
    l2, Dr2, t2 = extrapolate(files, os.path.join(str(tmp_path), "sys."))
    np.testing.assert_allclose(Dr2, Dreal, rtol=1e-6)
    assert os.path.exists(os.path.join(str(tmp_path), "sys.Dreal.dat")) and os.path.exists(os.path.join(str(tmp_path), "sys.t0.dat"))
    row = open(os.path.join(str(tmp_path), "sys.Dreal.dat")).read().splitlines()
    assert row[0] == "   index  bin-str  bin-end" and row[-1] == "=" * 10

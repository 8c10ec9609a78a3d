"""The C ABI: the shared library loads, exports every symbol include/mcdiff_b200.h declares,
and the ctypes structures have the layout the header promises.  No compute (no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from mcdiff_b200 import _capi
    return _capi.load()


def header_functions():
    text = open(os.path.join(ROOT, "include", "mcdiff_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcd_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    from mcdiff_b200 import _capi
    names = header_functions()
    assert set(names) == set(_capi.EXPORTS)
    for name in names:
        assert hasattr(lib, name), name


def test_version_and_error_strings(lib):
    assert lib.mcd_version() == 200          # MCD_VERSION of include/mcdiff_b200.h
    assert lib.mcd_error_string(0) == b"ok"
    assert lib.mcd_error_string(-1) == b"invalid argument"
    assert lib.mcd_error_string(-2) == b"unsupported configuration"


def test_struct_layouts():
    from mcdiff_b200 import _capi as K
    assert C.sizeof(K.ProblemDesc) == 6 * 4 + 5 * 8 + 3 * 8      # + double pull
    assert C.sizeof(K.McParams) == 3 * 8 + 6 * 4 + 4 * 8
    assert C.sizeof(K.ChainIO) == 15 * 8
    assert K.McParams.seed.offset == 48 and K.McParams.method.offset == 40
    assert C.sizeof(K.RadProblemDesc) == 6 * 4 + 7 * 8 + 8
    assert C.sizeof(K.RadMcParams) == 2 * 8 + 4 * 4 + 4 * 8 and C.sizeof(K.RadChainIO) == 11 * 8


def test_argument_errors_without_gpu(lib):
    # null handle / null output are argument errors, never a crash
    assert lib.mcd_create(0, None) == -1
    assert lib.mcd_destroy(None) == 0
    assert lib.mcd_problem_destroy(None) == 0
    assert lib.mcd_rad_problem_destroy(None) == 0
    assert lib.mcd_rad_problem_create(None, None, None, None) == -1
    assert lib.mcd_launch_count(None) == 0
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert lib.mcd_create(0, C.byref(h)) == -4          # MCD_E_NODEVICE


def test_no_cpu_fallback():
    """Without a GPU the product refuses to compute instead of silently using the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mcdiff_b200 import _capi
    from mcdiff_b200.engine import Engine
    with pytest.raises(_capi.McdError):
        Engine(0)
    import mcdiff_b200
    import sys
    assert not any(m.startswith("oracle") for m in sys.modules if "mcdiff_oracle" in m and "mcdiff_b200" in m)
    for mod in ("engine", "MCState", "mc", "log", "model", "transitions", "outreading", "parallel", "cli"):
        src = open(os.path.join(ROOT, "mcdiff_b200", mod + ".py")).read()
        assert "oracle" not in src, mod

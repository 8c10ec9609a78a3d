"""lib/b200.py -- the file a maintainer adds to annekegh/mcdiff.

ctypes binding of libmcdiff_b200.so for the likelihood seam of the reference: an object with the signature of
`mcdiff.utils.log_like_lag` (lib/utils.py:355-368) that `MCState` can call instead (lib/MCState.py:110,209,247,282).
It depends only on ctypes, numpy and torch (device buffers), NOT on the mcdiff_b200 python package.
tests/test_gpu_reference_patch.py installs it into the unmodified reference and checks K1-K4.
"""
import ctypes as C
import os

import numpy as np
import torch

_lib = C.CDLL(os.environ.get("MCDIFF_B200_LIB", "libmcdiff_b200.so"))


class _Desc(C.Structure):                       # mcd_problem_desc (include/mcdiff_b200.h)
    _fields_ = [("n", C.c_int32), ("pbc", C.c_int32), ("nlag", C.c_int32), ("ncosF", C.c_int32),
                ("ncosD", C.c_int32), ("use_pull", C.c_int32), ("lagtimes", C.c_void_p),
                ("counts", C.c_void_p), ("v_basis", C.c_void_p), ("w_basis", C.c_void_p),
                ("string_vecs", C.c_void_p), ("spring_k", C.c_double), ("clip", C.c_double),
                ("pull", C.c_double)]


_lib.mcd_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
_lib.mcd_problem_create.argtypes = [C.c_void_p, C.POINTER(_Desc), C.c_void_p, C.POINTER(C.c_void_p)]
_lib.mcd_loglike_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 7
_lib.mcd_problem_destroy.argtypes = [C.c_void_p]
_lib.mcd_destroy.argtypes = [C.c_void_p]


class B200Likelihood:
    """Drop-in for mcdiff.utils.log_like_lag on one GPU."""

    def __init__(self, data, pbc, device=0, pull=None):      # pull: dF per bin (MCState.pull) or None
        self.dev = torch.device("cuda", device)
        self.h = C.c_void_p()
        assert _lib.mcd_create(device, C.byref(self.h)) == 0
        self.counts = torch.as_tensor(np.asarray(data.list_trans), dtype=torch.int32, device=self.dev).contiguous()
        self.lags_host = np.asarray(data.list_lt, dtype=np.float64)
        self.lags = torch.as_tensor(self.lags_host, device=self.dev)
        d = _Desc(n=data.dim_trans, pbc=int(pbc), nlag=data.dim_lt, lagtimes=self.lags.data_ptr(),
                  counts=self.counts.data_ptr(), spring_k=-1.0, clip=1e-10,
                  use_pull=int(pull is not None), pull=float(pull or 0.0))
        self.pull = pull
        self.p = C.c_void_p()
        torch.cuda.synchronize(self.dev)
        assert _lib.mcd_problem_create(self.h, C.byref(d), None, C.byref(self.p)) == 0

    def __call__(self, num_bin, num_lag, v, w, lagtimes, transition, pbc, pull):
        assert pull == self.pull                               # the problem object was created for this force
        dv = torch.as_tensor(np.asarray(v, dtype=np.float64)[None, :], device=self.dev)
        dw = torch.as_tensor(np.asarray(w, dtype=np.float64)[None, :], device=self.dev)
        # per-profile lag times (time-offset trial, lib/MCState.py:207-210 = lag times + one offset): method 2
        # (SQUARING) walks the shifted LagPlan (MCD_ST_BAD_LAGS if they are not a pure offset);
        # NULL = the problem's own lag times, method 0 (AUTO)
        shifted = not np.array_equal(np.asarray(lagtimes, dtype=np.float64), self.lags_host)
        dl = torch.as_tensor(np.asarray(lagtimes, dtype=np.float64)[None, :], device=self.dev) if shifted else None
        out = torch.empty(1, dtype=torch.float64, device=self.dev)
        torch.cuda.synchronize(self.dev)                       # the library runs on the NULL stream here
        rc = _lib.mcd_loglike_batch(self.h, self.p, 2 if shifted else 0, 1, dv.data_ptr(), dw.data_ptr(),
                                    dl.data_ptr() if shifted else None, out.data_ptr(), None, None, None)
        assert rc == 0
        return np.float64(out.item())

    def close(self):
        if self.p:
            _lib.mcd_problem_destroy(self.p)
            _lib.mcd_destroy(self.h)
            self.p = None

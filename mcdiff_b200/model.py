"""Profile models: which numbers parametrise F(z) (v, kBT) and ln D(z) (w, in dx^2/dt).

Same class names, attributes and basis definitions as the reference (lib/model.py:29-363);
the host builds the basis matrices once, the kernel evaluates profile = basis . coeff
(lib/model.py:128-130) for every proposal.

    v[i]    free energy of bin i                         [kBT]
    w[i]    ln D between bin i and i+1                   [ln(dx^2/dt)], wunit = ln(dx^2/dt)
    wrad[i] ln D_radial in bin i                         [ln(dr^2/dt)], wradunit = ln(dr^2/dt)
"""
from __future__ import annotations

import copy

import numpy as np


class Model:
    """Per-bin model (lib/model.py:29-74)."""

    ncosF = 0   # the reference's plain Model lacks these (run-mcdiff --model Model crashes on it,
    ncosD = 0   # SURVEY.md 7.3); here the per-bin model simply has no basis functions.
    ncosP = 0

    def __init__(self, trans, D0):
        self.edges = trans.edges
        self.count = trans.count
        if trans.count == "pbc":
            self.nbins = len(self.edges) - 1
            self.dim_w = self.nbins
        elif trans.count == "cut":
            self.nbins = len(self.edges) - 1
            self.dim_w = self.nbins - 1
        elif trans.count == "pbccut":
            self.nbins = len(self.edges)
            self.dim_w = self.nbins
        else:
            raise ValueError("unknown count method %r" % trans.count)
        if trans.dim_trans != self.nbins:
            raise AssertionError("transition matrix dimension does not match the bin edges")
        self.dim_trans = self.nbins
        self.dim_v = self.nbins
        self.pbc = "pbc" in trans.count
        self.list_lt = copy.deepcopy(trans.list_lt)   # "effective" lag times (time-offset moves)
        self.init_model(D0)

    def init_model(self, D0):
        self.v = np.zeros(self.dim_v, dtype=np.float64)
        self.w = np.zeros(self.dim_w, dtype=np.float64)
        dt = 1.0                                        # ps (hard-coded in the reference, lib/model.py:66)
        dx = self.edges[1] - self.edges[0]
        self.vunit = 1.0
        self.wunit = np.log(dx ** 2 / dt)
        self.w += np.log(D0) - self.wunit
        self.timezero = 0.0

    # the kernel-facing view -------------------------------------------------------------
    @property
    def v_param(self):
        return self.v_coeff if self.ncosF > 0 else self.v

    @property
    def w_param(self):
        return self.w_coeff if self.ncosD > 0 else self.w


class CosinusModel(Model):
    """Cosine basis (lib/model.py:77-149)."""

    def __init__(self, trans, D0, ncosF, ncosD, ncosP=0):
        if ncosF < 0 or ncosD < 0:
            raise AssertionError("number of basis functions must be >= 0")
        self.ncosF, self.ncosD, self.ncosP = ncosF, ncosD, ncosP
        Model.__init__(self, trans, D0)
        if self.ncosF > 0:
            self.init_model_cosF()
        if self.ncosD > 0:
            self.init_model_cosD(D0)

    def create_basis_center(self, dim, ncos):
        x = np.arange(dim)
        return np.array([np.cos(2 * k * np.pi * (x + 0.5) / dim) / (k + 1) for k in range(ncos)]).transpose()

    def create_basis_border(self, dim, dim2, ncos):
        x = np.arange(dim)
        return np.array([np.cos(2 * k * np.pi * (x + 1.0) / dim2) / (k + 1) for k in range(ncos)]).transpose()

    def init_model_cosF(self):
        self.v_coeff = np.zeros(self.ncosF, dtype=np.float64)
        self.v_basis = self.create_basis_center(self.dim_v, self.ncosF)
        self.update_v()

    def init_model_cosD(self, D0):
        self.w_coeff = np.zeros(self.ncosD, dtype=np.float64)
        self.w_coeff[0] += np.log(D0) - self.wunit
        self.w_basis = self.create_basis_border(self.dim_w, self.dim_v, self.ncosD)   # period = dim_v
        self.update_w()

    def calc_profile(self, coeff, basis):
        return np.sum(coeff * basis, axis=1)

    def update_v(self, coeff=None):
        if coeff is not None:
            if len(coeff) != len(self.v_coeff):
                raise AssertionError("wrong number of coefficients")
            self.v_coeff = coeff
        self.v = self.calc_profile(self.v_coeff, self.v_basis)

    def update_w(self, coeff=None):
        if coeff is not None:
            if len(coeff) != len(self.w_coeff):
                raise AssertionError("wrong number of coefficients")
            self.w_coeff = coeff
        self.w = self.calc_profile(self.w_coeff, self.w_basis)


class SinusCosinusModel(CosinusModel):
    """cos terms k = 0..(ncos-1)/2 followed by sin terms k = 1..(ncos-1)/2 (lib/model.py:151-202)."""

    def __init__(self, trans, D0, ncosF, ncosD, ncosP=0):
        if ncosF % 2 != 1 or ncosD % 2 != 1:
            raise AssertionError("SinusCosinusModel needs odd numbers of basis functions")
        CosinusModel.__init__(self, trans, D0, ncosF, ncosD, ncosP)

    def create_basis_center(self, dim, ncos):
        x = np.arange(dim)
        h = (ncos + 1) // 2
        cols = [np.cos(2 * k * np.pi * (x + 0.5) / dim) / (k + 1) for k in range(h)]
        cols += [np.sin(2 * k * np.pi * (x + 0.5) / dim) / (k + 1) for k in range(1, h)]
        return np.array(cols).transpose()

    def create_basis_border(self, dim, dim2, ncos):
        x = np.arange(dim)
        h = (ncos + 1) // 2
        cols = [np.cos(2 * k * np.pi * (x + 1.0) / dim2) / (k + 1) for k in range(h)]
        cols += [np.sin(2 * k * np.pi * (x + 1.0) / dim2) / (k + 1) for k in range(1, h)]
        return np.array(cols).transpose()


class StepModel(CosinusModel):
    """Nested symmetric box functions (lib/model.py:205-279)."""

    def __init__(self, trans, D0, ncosF, ncosD, ncosP=0):
        if ncosF < 0 or ncosD < 0:
            raise AssertionError("number of basis functions must be >= 0")
        self.ncosF, self.ncosD, self.ncosP = ncosF, ncosD, ncosP
        Model.__init__(self, trans, D0)
        if self.ncosF > 0:
            self.init_model_F()
        if self.ncosD > 0:
            self.init_model_D(D0)

    def _x0(self, ncos):
        dx = self.dim_v / 2.0 / ncos
        return np.arange(0, ncos * dx, dx)

    def init_model_F(self):
        self.v_coeff = np.zeros(self.ncosF, dtype=np.float64)
        self.v_x0 = self._x0(self.ncosF)
        self.v_basis = self.create_step_center(self.dim_v, self.v_x0)
        self.update_v()

    def init_model_D(self, D0):
        self.w_coeff = np.zeros(self.ncosD, dtype=np.float64)
        self.w_x0 = self._x0(self.ncosD)
        self.w_basis = self.create_step_border(self.dim_w, self.dim_v, self.w_x0)
        self._init_w_coeff(D0)
        self.update_w()

    def _init_w_coeff(self, D0):
        self.w_coeff += np.log(D0) - self.wunit          # every height, lib/model.py:235

    def create_step_center(self, dim, allx0):
        x = np.arange(dim) + 0.5
        return np.array([np.where((x >= x0) & (x <= dim - x0), 1.0, 0.0) for x0 in allx0]).transpose()

    def create_step_border(self, dim, dim2, allx0):
        x = np.arange(dim) + 1.0
        return np.array([np.where((x >= x0) & (x <= dim2 - x0), 1.0, 0.0) for x0 in allx0]).transpose()


class OneStepModel(StepModel):
    """One-sided steps (lib/model.py:281-310)."""

    def _x0(self, ncos):
        dx = float(self.dim_v) / ncos
        return np.arange(0, ncos * dx, dx)

    def _init_w_coeff(self, D0):
        self.w_coeff[0] += np.log(D0) - self.wunit       # first height only, lib/model.py:300

    def create_step_center(self, dim, allx0):
        x = np.arange(dim) + 0.5
        return np.array([np.where(x >= x0, 1.0, 0.0) for x0 in allx0]).transpose()

    def create_step_border(self, dim, dim2, allx0):
        x = np.arange(dim) + 1.0
        return np.array([np.where(x >= x0, 1.0, 0.0) for x0 in allx0]).transpose()


class RadModel(CosinusModel):
    """Adds the radial diffusion profile wrad (lib/model.py:313-336)."""

    ncosDrad = 0

    def __init__(self, trans, D0, ncosF, ncosD, ncosP=0):
        CosinusModel.__init__(self, trans, D0, ncosF, ncosD, ncosP)
        self.redges = trans.redges
        self.dim_rad = trans.dim_rad
        self.init_model_rad(D0)

    def init_model_rad(self, D0):
        self.wrad = np.zeros(self.dim_v, dtype=np.float64)
        self.dim_wrad = len(self.wrad)
        dr = self.redges[1] - self.redges[0]
        unit = dr ** 2 / 1.0
        self.wrad += np.log(D0 / unit)
        self.wradunit = np.log(unit)


class RadCosinusModel(RadModel):
    """Cosine basis for wrad (lib/model.py:338-363)."""

    def __init__(self, trans, D0, ncosF, ncosD, ncosP, ncosDrad):
        RadModel.__init__(self, trans, D0, ncosF, ncosD, ncosP)
        if ncosDrad < 0:
            raise AssertionError("ncosDrad must be >= 0")
        self.ncosDrad = ncosDrad
        if ncosDrad > 0:
            self.wrad_coeff = np.zeros(ncosDrad, dtype=np.float64)
            self.wrad_coeff[0] = np.log(D0) - self.wradunit
            self.wrad_basis = self.create_basis_center(self.dim_v, ncosDrad)
            self.update_wrad()

    def update_wrad(self, coeff=None):
        if coeff is not None:
            if len(coeff) != len(self.wrad_coeff):
                raise AssertionError("wrong number of coefficients")
            self.wrad_coeff = coeff
        self.wrad = self.calc_profile(self.wrad_coeff, self.wrad_basis)


MODELS_1D = {"CosinusModel": CosinusModel, "SinusCosinusModel": SinusCosinusModel, "StepModel": StepModel,
             "OneStepModel": OneStepModel}

"""Transition-count input: same text format, same attribute names as the reference's
`Transitions` / `RadTransitions` (lib/transitions.py:22-70, 200-253; parsing rules of
lib/reading.py:14-146).  Host-side only: the arrays built here are what the engine uploads.

File format (written by the reference's `write_Tmat_square` / `write_Tmat_cube`,
lib/tools/extract.py:19-62): header lines `#lt`, `#count pbc|cut`, `#dt`, `#dn`, `#edges ...`,
`#redges ...`; then one matrix row per line; cube files separate radial slabs by a line `-`.
"""
from __future__ import annotations

import numpy as np


def read_transition_header(filename):
    """Header dict with the reference's keys and consistency rules (lib/reading.py:43-105)."""
    header = {}
    with open(filename) as f:
        for line in f:
            if not line.startswith("#"):
                continue
            wds = line.split()
            key = wds[0]
            if key == "#edges":
                e = np.array([float(x) for x in wds[1:]])
                header["edges"] = e
                header["dz"] = (e[-1] - e[0]) / float(len(e) - 1)
            elif key == "#redges":
                header["redges"] = np.array([float(x) for x in wds[1:]])
            elif key == "#dt":
                header["dt"] = float(wds[1])
            elif key == "#dn":
                header["dn"] = int(wds[1])
            elif key == "#lt":
                header["lt"] = float(wds[1])
            elif key == "#count":
                header["count"] = wds[1]
            elif key == "#file":
                header["file"] = wds[1]
            elif key == "#abc":
                header.update(a=float(wds[1]), b=float(wds[2]), c=float(wds[3]))
    has = lambda k: k in header  # noqa: E731
    if has("dt") and has("dn"):
        if has("lt"):
            if abs(header["lt"] - header["dn"] * header["dt"]) >= 1e-5:
                raise AssertionError("lt != dn*dt in %s" % filename)
        else:
            header["lt"] = header["dn"] * header["dt"]
    elif has("dt") and has("lt"):
        header.setdefault("dn", header["lt"] * header["dt"])      # sic, lib/reading.py:88
    elif has("dn") and has("lt"):
        header.setdefault("dt", header["lt"] * header["dn"])      # sic, lib/reading.py:91
    else:
        raise ValueError("can not construct lag time lt")
    if not has("count"):
        raise ValueError("count should be specified in transition matrix file")
    if header["count"] not in ("pbc", "cut"):
        raise AssertionError("unknown count method %r" % header["count"])
    return header


def guess_dim_transition_square(filename):
    with open(filename) as f:
        for line in f:
            if not line.startswith("#"):
                return len(line.split())
    raise ValueError("no data in %s" % filename)


def read_transition_square(filename, dim_trans):
    """Integer n x n matrix, one row per line (lib/reading.py:108-124)."""
    rows = []
    with open(filename) as f:
        for line in f:
            if line.startswith("#"):
                continue
            wds = line.split()
            if len(wds) != dim_trans:
                raise AssertionError("row of %d entries, expected %d" % (len(wds), dim_trans))
            rows.append([int(x) for x in wds])
    if len(rows) != dim_trans:
        raise ValueError("wrong number of entries in %s" % filename)
    return np.array(rows, dtype=np.int64)


def guess_dim_transition_cube(filename):
    dim_rad, row, lenrow = 0, 0, 0
    with open(filename) as f:
        for line in f:
            if line.startswith("-"):
                dim_rad += 1
                row = 0
            elif not line.startswith("#"):
                lenrow = len(line.split())
                row += 1
    return dim_rad, lenrow


def read_transition_cube(filename, dim_rad, dim_trans):
    """float [dim_rad][n][n] (lib/reading.py:126-146)."""
    out = np.zeros((dim_rad, dim_trans, dim_trans))
    k = row = 0
    with open(filename) as f:
        for line in f:
            if line.startswith("-"):
                if row != dim_trans:
                    raise ValueError("wrong number of entries in %s" % filename)
                k += 1
                row = 0
            elif not line.startswith("#"):
                wds = line.split()
                if len(wds) != dim_trans:
                    raise AssertionError("bad row length")
                out[k, row, :] = [int(x) for x in wds]
                row += 1
    return out


def reduce_Tmat(dim_trans, header, transmatrix):
    """--reduction: drop leading/trailing bins with an all-zero row or column; the result is a
    `cut` matrix (lib/transitions.py:72-91)."""
    sel = [i for i in range(len(transmatrix)) if transmatrix[i, :].sum() != 0 and transmatrix[:, i].sum() != 0]
    sel = list(range(min(sel), max(sel) + 1))
    redux = len(transmatrix) - len(sel)
    # like the reference, the result is flagged "cut" even when nothing had to be removed
    # (the "nothing happened" test of lib/transitions.py:79 can never be true)
    trans = np.array(transmatrix[np.ix_(sel, sel)], dtype=np.float64)
    header = dict(header)
    header["count"] = "cut"
    header["edges"] = header["edges"][sel + [sel[-1] + 1]]
    return dim_trans - redux, header, trans


class Transitions:
    """Count matrices of several lag times (same public attributes as the reference class)."""

    def __init__(self, list_filenames, reduction=False):
        self.dim_lt = len(list_filenames)
        if self.dim_lt <= 0:
            raise AssertionError("need at least one transition file")
        self.list_filenames = list(list_filenames)
        lt, dt, dn, tr = [], [], [], []
        self.started = False
        for fn in self.list_filenames:
            dim = guess_dim_transition_square(fn)
            header = read_transition_header(fn)
            mat = read_transition_square(fn, dim)
            if reduction:
                dim, header, mat = reduce_Tmat(dim, header, mat)
            self._settings(header, dim)
            lt.append(header["lt"]); dt.append(header["dt"]); dn.append(header["dn"]); tr.append(mat)
        self.list_lt = np.array(lt)
        self.list_dt = np.array(dt)
        self.list_dn = np.array(dn)
        self.list_trans = np.array(tr)
        self.min_lt = min(self.list_lt)

    def _settings(self, header, dim):
        if not self.started:
            self.started = True
            self.count = header["count"]
            self.dim_trans = dim
            if "edges" in header:
                self.edges = header["edges"]
                self.dz = header["dz"]
            else:
                self.edges = np.arange(dim + 1.0)
                self.dz = 1.0
        else:
            if self.count != header["count"] or self.dim_trans != dim:
                raise AssertionError("transition files are not compatible")
            if "edges" in header:
                if not (self.edges == header["edges"]).all() or self.dz != header["dz"]:
                    raise AssertionError("transition files have different edges")

    @classmethod
    def from_arrays(cls, counts, lagtimes, edges, count="pbc"):
        """In-memory construction (synthetic workloads, tests)."""
        self = cls.__new__(cls)
        self.list_filenames = []
        self.list_trans = np.asarray(counts)
        self.dim_lt = self.list_trans.shape[0]
        self.list_lt = np.asarray(lagtimes, dtype=np.float64)
        self.list_dt = np.ones(self.dim_lt)
        self.list_dn = np.asarray(self.list_lt, dtype=np.int64)
        self.min_lt = min(self.list_lt)
        self.count = count
        self.dim_trans = self.list_trans.shape[-1]
        self.edges = np.asarray(edges, dtype=np.float64)
        self.dz = (self.edges[-1] - self.edges[0]) / float(len(self.edges) - 1)
        self.started = True
        return self


class RadTransitions:
    """Radial count cubes [dim_rad][n][n] per lag time (lib/transitions.py:200-253)."""

    def __init__(self, list_filenames):
        self.dim_lt = len(list_filenames)
        if self.dim_lt <= 0:
            raise AssertionError("need at least one transition file")
        self.list_filenames = list(list_filenames)
        lt, dt, dn, tr = [], [], [], []
        self.started = False
        for fn in self.list_filenames:
            dim_rad, dim = guess_dim_transition_cube(fn)
            header = read_transition_header(fn)
            cube = read_transition_cube(fn, dim_rad, dim)
            if not self.started:
                self.started = True
                self.count = header["count"]
                self.dim_trans, self.dim_rad = dim, dim_rad
                if "edges" in header:
                    self.edges, self.dz = header["edges"], header["dz"]
                else:
                    self.edges, self.dz = np.arange(dim + 1.0), 1.0
                if "redges" not in header:
                    raise ValueError("radial edges (#redges) are required in %s" % fn)
                self.redges = header["redges"]
            else:
                if (self.count, self.dim_trans, self.dim_rad) != (header["count"], dim, dim_rad):
                    raise AssertionError("transition files are not compatible")
                if "edges" in header and not (self.edges == header["edges"]).all():
                    raise AssertionError("different edges")
                if "redges" in header and not (self.redges == header["redges"]).all():
                    raise AssertionError("different redges")
            lt.append(header["lt"]); dt.append(header["dt"]); dn.append(header["dn"]); tr.append(cube)
        self.list_lt = np.array(lt)
        self.list_dt = np.array(dt)
        self.list_dn = np.array(dn)
        self.list_trans = np.array(tr)
        self.min_lt = min(self.list_lt)


def write_Tmat_square(A, filename, lt, count, edges=None, dt=None, dn=None):
    """Same text layout as the reference writer (lib/tools/extract.py:19-36)."""
    with open(filename, "w") as f:
        f.write("#lt    {}\n".format(lt))
        f.write("#count {}\n".format(count))
        if dt is not None:
            f.write("#dt    {}\n".format(dt))
        if dn is not None:
            f.write("#dn    {}\n".format(dn))
        if edges is not None:
            f.write("#edges  " + " ".join(str(b) for b in edges) + "\n")
        for row in np.asarray(A):
            f.write(" ".join(str(x) for x in row) + "\n")


def write_Tmat_cube(B, filename, lt, count, edges=None, redges=None, dt=None, dn=None):
    """lib/tools/extract.py:38-62."""
    B = np.asarray(B)
    with open(filename, "w") as f:
        f.write("#lt    {}\n".format(lt))
        f.write("#count {}\n".format(count))
        if dt is not None:
            f.write("#dt    {}\n".format(dt))
        if dn is not None:
            f.write("#dn    {}\n".format(dn))
        if edges is not None:
            f.write("#edges  " + " ".join(str(b) for b in edges) + "\n")
        if redges is not None:
            f.write("#redges  " + " ".join(str(b) for b in redges) + "\n")
        for slab in B:
            for row in slab:
                f.write(" ".join(str(x) for x in row) + "\n")
            f.write("-\n")

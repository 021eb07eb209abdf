"""A single-rank numpy stand-in for the slice of the ``pmesh.pm`` API that ``hymd/field.py`` uses.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Purpose: run the REFERENCE's own, unmodified ``hymd/field.py`` / ``hymd/pressure.py`` source (loaded
from ``/root/reference`` by ``tests/golden/ref_loader.py``) on top of ``oracle/pm_oracle.py``'s
restatement of the pmesh primitives, and so turn ``oracle/field_oracle.py`` from "a restatement of
field.py" into "a restatement checked against field.py itself": after this, only the pmesh primitives
(CIC window, FFT normalisation, wave-vector convention, ``normp``) remain restated, and those are
pinned by the reference's energy KATs (``tests/test_oracle_reference_kats.py``) and the analytic
results of ``tests/test_oracle_analytic.py``.

API covered, with the reference call sites that use it:

* ``ParticleMesh(Nmesh, BoxSize=, dtype=, comm=)``           ``field.py:45-47``
* ``pm.create("real"|"complex", value=0.0)``                 ``field.py:48-137``, ``main.py:488-498``
* ``pm.paint(pos, layout=, mass=, out=)``                    ``field.py:363, 574``
* ``pm.decompose(pos, smoothing=)`` -> layout with ``exchange`` / ``get_exchange_cost``
                                                             ``main.py:977-980``, ``field.py:1165-1178``
* ``RealField``: ndarray arithmetic, ``.value``, ``.r2c(out=)``, ``.readout(pos, layout=)``, ``.csum()``
                                                             ``field.py:200, 366, 575-576, 693``
* ``ComplexField``: ``.value``, ``.apply(func, out=Ellipsis|field)`` with ``k[d]`` and
  ``k.normp(p=2, zeromode=1)``, ``.c2r(out=)``               ``field.py:369-397, 577-616``
"""
from __future__ import annotations

import numpy as np

from . import pm_oracle as pmo


class _K(list):
    """What pmesh hands to an ``apply`` callback as ``k``: the three broadcastable wave-number
    arrays plus ``normp``."""

    def normp(self, p=2, zeromode=None):
        out = np.array(sum(np.abs(ki) ** p for ki in self), copy=True)
        if zeromode is not None:
            out[out == 0] = zeromode
        return out


class _Field(np.ndarray):
    pm = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self.pm = getattr(obj, "pm", None)

    @property
    def value(self):
        return self.view(np.ndarray)

    def copy(self, order="C"):
        out = np.ndarray.copy(self, order).view(type(self))
        out.pm = self.pm
        return out


class RealField(_Field):
    def r2c(self, out=None):
        res = pmo.r2c(np.ascontiguousarray(self.value))
        if out is None:
            out = self.pm.create("complex")
        out.value[...] = res
        return out

    def readout(self, pos, layout=None, out=None):
        return pmo.cic_readout(np.ascontiguousarray(self.value), np.asarray(pos), self.pm.BoxSize,
                               use_c=False)

    def csum(self):
        return pmo.csum(self.value)


class ComplexField(_Field):
    def apply(self, func, kind="wavenumber", out=None):
        if out is Ellipsis:
            out = self
        elif out is None:
            out = self.pm.create("complex")
        res = func(self.pm._k, self.value)
        out.value[...] = res
        return out

    def c2r(self, out=None):
        res = pmo.c2r(self.value, self.pm.Nmesh)
        if out is None:
            out = self.pm.create("real")
        out.value[...] = res
        return out


class Layout:
    """Single rank: every particle stays."""

    def __init__(self, n):
        self.n = n

    def exchange(self, *arrays):
        if len(arrays) == 1:
            return arrays[0]
        return list(arrays)

    def get_exchange_cost(self):
        return np.zeros(1, dtype=np.int64)


class ParticleMesh:
    def __init__(self, Nmesh, BoxSize=1.0, dtype="f8", comm=None, np=None):
        import numpy as _np
        self.Nmesh = _np.array(pmo.mesh_tuple(Nmesh), dtype=_np.int64)
        self.BoxSize = _np.empty(3, dtype=_np.float64)
        self.BoxSize[:] = BoxSize
        self.dtype = _np.dtype(dtype)
        self.comm = comm
        self.np = _np.array([1, 1])
        self._k = _K(pmo.kgrid(tuple(self.Nmesh), self.BoxSize, self.dtype))

    def create(self, type="real", value=None):
        nx, ny, nz = (int(v) for v in self.Nmesh)
        if type == "real":
            f = np.zeros((nx, ny, nz), dtype=self.dtype).view(RealField)
        elif type == "complex":
            cdt = np.complex128 if self.dtype == np.float64 else np.complex64
            f = np.zeros((nx, ny, nz // 2 + 1), dtype=cdt).view(ComplexField)
        else:
            raise ValueError(type)
        f.pm = self
        if value is not None:
            f.value[...] = value
        return f

    def decompose(self, pos, smoothing=None):
        return Layout(len(pos))

    def paint(self, pos, mass=1.0, layout=None, out=None, hold=False):
        res = pmo.cic_paint(np.asarray(pos), mass, tuple(self.Nmesh), self.BoxSize, self.dtype,
                            use_c=False)
        if out is None:
            out = self.create("real")
        if hold:
            out.value[...] += res
        else:
            out.value[...] = res
        return out

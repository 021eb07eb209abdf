"""Restatement of the pmesh primitives on HyMD's field-force path (CPU, numpy).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

pmesh itself (third party, unpinned, not vendored, not installed) cannot be
imported here, so this file restates the published algorithm of the handful of
primitives that ``hymd/field.py`` calls, anchored on the reference call sites:

=====================  ==========================================  ============
primitive              reference call sites                        here
=====================  ==========================================  ============
``pm.paint``           ``field.py:363, 574``                       cic_paint
``RealField.readout``  ``field.py:200, 402``                       cic_readout
``RealField.r2c``      ``field.py:366, 576, 584``                  r2c
``ComplexField.c2r``   ``field.py:377, 397, 578, 613, 616``        c2r
``ComplexField.apply`` ``field.py:369, 375, 396, 577, 585, 612``   wavevectors
``k.normp(2, zeromode=1)``  ``field.py:373, 391``                  knorm2_zeromode1
``RealField.csum``     ``field.py:693, 699``                       csum
=====================  ==========================================  ============

Conventions (each is exercised by the known-answer tests):

* mesh vertex ``j`` sits at ``j * L/N`` (vertex centred); cloud-in-cell:
  ``x = r*N/L, c = floor(x), d = x - c``; the 8 vertices ``(c + delta) mod N``
  get ``mass * prod(d if delta else 1-d)``; ``paint(out=...)`` overwrites.
* ``r2c`` carries the ``1/M`` normalisation, ``c2r`` none: ``c2r(r2c(x)) == x``.
* wave vectors follow ``numpy.fft.fftfreq``: ``k_i(n) = 2 pi n~/L_i`` with
  ``n~ = n`` for ``n < N/2`` else ``n - N``, on all three axes including the
  halved (last) one, so that an even-N Nyquist index carries ``-pi N/L``.
* ``c2r`` has FFTW / ``irfftn`` semantics on the stored half spectrum (complex
  inverse transforms over x and y, then a 1-D c2r over z that ignores the
  imaginary part of the ``k_z = 0`` and ``k_z = N_z/2`` elements).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import scipy.fft as _fft

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB_PATH = os.path.join(_HERE, "_build", "libcic_oracle.so")
_clib = None


def _load_clib():
    """Optional C restatement of the CIC loops (``oracle/cic_oracle.c``)."""
    global _clib
    if _clib is None and os.path.exists(_CLIB_PATH):
        lib = ctypes.CDLL(_CLIB_PATH)
        for name in ("cic_paint_f32", "cic_paint_f64", "cic_readout_f32", "cic_readout_f64",
                     "cic_paint_mt_f32", "cic_paint_mt_f64", "cic_readout_mt_f32",
                     "cic_readout_mt_f64"):
            getattr(lib, name).restype = None
        _clib = lib
    return _clib


def mesh_tuple(mesh_size):
    """``np.full(3, config.mesh_size)`` (``field.py:571``)."""
    m = np.full(3, mesh_size).astype(np.int64)
    return int(m[0]), int(m[1]), int(m[2])


def wavevectors(mesh, box, dtype=np.float64):
    """Per-axis angular wave numbers ``(kx, ky, kz_half)`` of the stored spectrum."""
    nx, ny, nz = mesh_tuple(mesh)
    box = np.asarray(box, dtype=np.float64)
    kx = 2.0 * np.pi * np.fft.fftfreq(nx, d=box[0] / nx)
    ky = 2.0 * np.pi * np.fft.fftfreq(ny, d=box[1] / ny)
    kz = (2.0 * np.pi * np.fft.fftfreq(nz, d=box[2] / nz))[: nz // 2 + 1]
    return kx.astype(dtype), ky.astype(dtype), kz.astype(dtype)


def kgrid(mesh, box, dtype=np.float64):
    """Broadcastable ``[kx[:,None,None], ky[None,:,None], kz[None,None,:]]`` (what
    pmesh hands to an ``apply`` callback as ``k``)."""
    kx, ky, kz = wavevectors(mesh, box, dtype)
    return [kx[:, None, None], ky[None, :, None], kz[None, None, :]]


def knorm2_zeromode1(k):
    """``k.normp(p=2, zeromode=1)``: |k|^2 with the origin replaced by 1."""
    k2 = k[0] ** 2 + k[1] ** 2 + k[2] ** 2
    k2 = np.array(k2, copy=True)
    k2[0, 0, 0] = 1.0
    return k2


def _cic_indices(pos, mesh, box, dtype):
    nx, ny, nz = mesh_tuple(mesh)
    n = np.array([nx, ny, nz], dtype=np.int64)
    scale = (n / np.asarray(box, dtype=np.float64)).astype(dtype)
    x = np.asarray(pos, dtype=dtype) * scale[None, :]
    c = np.floor(x)
    d = (x - c).astype(dtype)
    c = c.astype(np.int64)
    return n, c, d


def cic_paint(pos, mass, mesh, box, dtype=np.float64, use_c=True):
    """Cloud-in-cell deposit of ``mass`` (scalar or per-particle) -> (Nx,Ny,Nz)."""
    dtype = np.dtype(dtype)
    nx, ny, nz = mesh_tuple(mesh)
    pos = np.ascontiguousarray(pos, dtype=dtype).reshape(-1, 3)
    npart = pos.shape[0]
    m = np.ascontiguousarray(np.broadcast_to(np.asarray(mass, dtype=dtype), (npart,)))
    out = np.zeros((nx, ny, nz), dtype=dtype)
    lib = _load_clib() if use_c else None
    if lib is not None:
        fn = lib.cic_paint_f64 if dtype == np.float64 else lib.cic_paint_f32
        b = np.asarray(box, dtype=np.float64)
        fn(pos.ctypes.data_as(ctypes.c_void_p), m.ctypes.data_as(ctypes.c_void_p),
           ctypes.c_long(npart), ctypes.c_int(nx), ctypes.c_int(ny), ctypes.c_int(nz),
           ctypes.c_double(b[0]), ctypes.c_double(b[1]), ctypes.c_double(b[2]),
           out.ctypes.data_as(ctypes.c_void_p))
        return out
    n, c, d = _cic_indices(pos, mesh, box, dtype)
    one = dtype.type(1.0)
    flat = out.reshape(-1)
    for ax in (0, 1):
        for ay in (0, 1):
            for az in (0, 1):
                w = m * (d[:, 0] if ax else one - d[:, 0])
                w = w * (d[:, 1] if ay else one - d[:, 1])
                w = w * (d[:, 2] if az else one - d[:, 2])
                ix = np.mod(c[:, 0] + ax, n[0])
                iy = np.mod(c[:, 1] + ay, n[1])
                iz = np.mod(c[:, 2] + az, n[2])
                np.add.at(flat, (ix * n[1] + iy) * n[2] + iz, w.astype(dtype))
    return out


def cic_readout(field, pos, box, use_c=True):
    """Cloud-in-cell interpolation of ``field`` (Nx,Ny,Nz) at ``pos`` -> (N,)."""
    dtype = field.dtype
    mesh = field.shape
    pos = np.ascontiguousarray(pos, dtype=dtype).reshape(-1, 3)
    npart = pos.shape[0]
    lib = _load_clib() if use_c else None
    if lib is not None and field.flags.c_contiguous:
        out = np.empty(npart, dtype=dtype)
        fn = lib.cic_readout_f64 if dtype == np.float64 else lib.cic_readout_f32
        b = np.asarray(box, dtype=np.float64)
        fn(field.ctypes.data_as(ctypes.c_void_p), pos.ctypes.data_as(ctypes.c_void_p),
           ctypes.c_long(npart), ctypes.c_int(mesh[0]), ctypes.c_int(mesh[1]),
           ctypes.c_int(mesh[2]), ctypes.c_double(b[0]), ctypes.c_double(b[1]),
           ctypes.c_double(b[2]), out.ctypes.data_as(ctypes.c_void_p))
        return out
    n, c, d = _cic_indices(pos, mesh, box, dtype)
    one = dtype.type(1.0)
    out = np.zeros(npart, dtype=dtype)
    for ax in (0, 1):
        for ay in (0, 1):
            for az in (0, 1):
                w = (d[:, 0] if ax else one - d[:, 0])
                w = w * (d[:, 1] if ay else one - d[:, 1])
                w = w * (d[:, 2] if az else one - d[:, 2])
                ix = np.mod(c[:, 0] + ax, n[0])
                iy = np.mod(c[:, 1] + ay, n[1])
                iz = np.mod(c[:, 2] + az, n[2])
                out += w * field[ix, iy, iz]
    return out


def r2c(x, workers=1):
    """``RealField.r2c``: forward real FFT carrying the 1/M normalisation."""
    y = _fft.rfftn(x, workers=workers)
    y *= x.dtype.type(1.0 / x.size)
    return y


def c2r(y, mesh, workers=1):
    """``ComplexField.c2r``: unnormalised inverse of :func:`r2c` (irfftn semantics)."""
    nx, ny, nz = mesh_tuple(mesh)
    x = _fft.irfftn(y, s=(nx, ny, nz), workers=workers)
    x *= x.dtype.type(nx * ny * nz)
    return x


def csum(x):
    """``RealField.csum``: global sum (single rank here)."""
    return float(np.sum(x, dtype=np.float64))

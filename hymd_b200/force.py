"""Device versions of HyMD's intramolecular force kernels.

Same names, argument order and in-place output convention as the f2py kernels that ``hymd/main.py``
calls ``respa_inner`` times per outer step (``main.py:841-887``):

=============================  ==========================================================
this module                    reference
=============================  ==========================================================
``compute_bond_forces``        ``cbf``  ``hymd/compute_bond_forces.f90:1-61``
``compute_angle_forces``       ``caf``  ``hymd/compute_angle_forces.f90:1-93``
``compute_dihedral_forces``    ``cdf``  ``hymd/compute_dihedral_forces.f90:1-137`` (dtype 0, 1, 2;
                               dipoles and transfer matrices of ``dipole_reconstruction.f90``)
``dipole_forces_redistribution``  ``hymd/force.py:855-880``
=============================  ==========================================================

The index / parameter arrays are the ones ``prepare_bonds`` returns (``hymd/force.py:573-728``;
host numpy arrays).  They are uploaded once into a device-resident :class:`BondedTopology`
(``hymd_bonded_create``: per-particle term lists, see ``csrc/bonded.cuh``) that is cached on the
identity of the index arrays, so the per-step calls only launch kernels.  Positions and forces may
be torch CUDA tensors (no copies) or numpy arrays (copied in and out).  There is no CPU fallback.

Return values follow the reference (energy, pressure by-product).  With CUDA tensor inputs they are
0-dim / (3,) float64 DEVICE tensors, so the inner rRESPA loop never synchronizes with the host
(``main.py:964-970`` only reads the energies of the last inner step); with numpy inputs they are a
Python float and a numpy array exactly like the f2py kernels.
"""
from __future__ import annotations

import ctypes
import weakref

import numpy as np
import torch

from . import _lib

_I32P = ctypes.POINTER(ctypes.c_int32)
_F64P = ctypes.POINTER(ctypes.c_double)


def _i32(x):
    return np.ascontiguousarray(np.asarray(x), dtype=np.int32).reshape(-1)


def _f64(x):
    return np.ascontiguousarray(np.asarray(x), dtype=np.float64)


def _default_device():
    return f"cuda:{torch.cuda.current_device()}"


class BondedTopology:
    """Device-resident term lists of one rank's molecules (``hymd_bonded`` handle)."""

    def __init__(self, n_particles, bonds=None, angles=None, dihedrals=None, device=None):
        """``bonds = (a, b, r0, k)``, ``angles = (a, b, c, theta0, k)``,
        ``dihedrals = (a, b, c, d, coeff (D,6,5), dih_type[, last])``; any of them may be ``None``.
        ``last`` = ``bonds_4_last`` (1 = last dihedral of a backbone; only read for ``dih_type`` 1)."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.HymdError("hymd_b200.force needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else _default_device())
        self.n_particles = int(n_particles)
        empty_i, empty_f = np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.float64)
        a2, b2, r0, k2 = bonds if bonds is not None else (empty_i, empty_i, empty_f, empty_f)
        a3, b3, c3, t0, k3 = angles if angles is not None else (empty_i,) * 3 + (empty_f,) * 2
        last4 = None
        if dihedrals is not None:
            a4, b4, c4, d4, coeff, dtype4 = dihedrals[:6]
            if len(dihedrals) > 6 and dihedrals[6] is not None:
                last4 = _i32(dihedrals[6])
        else:
            a4 = b4 = c4 = d4 = dtype4 = empty_i
            coeff = np.zeros((0, 6, 5))
        keep = [_i32(a2), _i32(b2), _f64(r0).reshape(-1), _f64(k2).reshape(-1),
                _i32(a3), _i32(b3), _i32(c3), _f64(t0).reshape(-1), _f64(k3).reshape(-1),
                _i32(a4), _i32(b4), _i32(c4), _i32(d4), _f64(coeff).reshape(-1), _i32(dtype4)]
        self.n_terms = (len(keep[0]), len(keep[4]), len(keep[9]))
        if len(keep[13]) != 30 * self.n_terms[2]:
            raise ValueError("bonds_4_coeff must have shape (D, 6, 5) (prepare_bonds, force.py:678-690)")

        def ip(x):
            return x.ctypes.data_as(_I32P)

        def fp(x):
            return x.ctypes.data_as(_F64P)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.hymd_bonded_create(
                self.n_particles,
                self.n_terms[0], ip(keep[0]), ip(keep[1]), fp(keep[2]), fp(keep[3]),
                self.n_terms[1], ip(keep[4]), ip(keep[5]), ip(keep[6]), fp(keep[7]), fp(keep[8]),
                self.n_terms[2], ip(keep[9]), ip(keep[10]), ip(keep[11]), ip(keep[12]), fp(keep[13]),
                ip(keep[14]), ctypes.byref(handle)))
        self._h = handle
        self.n_cbt = int(np.count_nonzero(keep[14] == 1))
        if last4 is not None:
            if len(last4) != self.n_terms[2]:
                raise ValueError("bonds_4_last must have one entry per dihedral")
            _lib.check(self.lib.hymd_bonded_set_last(handle, ip(last4)))
        self._out = torch.zeros((3, 4), dtype=torch.float64, device=self.device)
        self._out12 = torch.zeros((3, 4), dtype=torch.float64, device=self.device)
        self._cta = None     # unknown until set (HYMD_B200_BONDED_CTA may have chosen at creation)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.hymd_bonded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_cta(self, enable=True):
        """Switch between per-particle (0, default) and CTA-cooperative term evaluation
        (``hymd_bonded_set_cta``; 1: indirect term lists, 2: inline records + positions staged in shared
        memory): same forces, 2-4x fewer term evaluations."""
        _lib.check(self.lib.hymd_bonded_set_cta(self._h, int(enable)))
        self._cta = int(enable)

    def set_math(self, f32=True):
        """Single-precision arithmetic for bonds and angles in the per-particle fused inner step of the
        fp32 build (``hymd_bonded_set_math``, ``csrc/bonded_f32.cuh``); the default is the Fortran's double
        arithmetic."""
        _lib.check(self.lib.hymd_bonded_set_math(self._h, 1 if f32 else 0))

    def launch_count(self):
        return int(self.lib.hymd_bonded_launch_count(self._h))

    def forces(self, kind, positions, box_size, out):
        """Launch the kernel for ``kind`` (2, 3 or 4 particles per term): ``out`` (n,3) device
        tensor is overwritten; returns the (4,) float64 device tensor {energy, pr_x, pr_y, pr_z}."""
        n = self.n_particles
        if tuple(positions.shape) != (n, 3) or tuple(out.shape) != (n, 3):
            raise ValueError(f"positions / forces must have shape ({n}, 3)")
        if positions.dtype != out.dtype or positions.dtype not in (torch.float32, torch.float64):
            raise ValueError("positions and forces must share dtype float32 or float64")
        box = (ctypes.c_double * 3)(*[float(b) for b in np.asarray(box_size).reshape(-1)[:3]])
        res = self._out[kind - 2]
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.hymd_bonded_forces(
            self._h, int(kind), _lib.F64 if positions.dtype == torch.float64 else _lib.F32,
            ctypes.c_void_p(positions.data_ptr()), box, ctypes.c_void_p(out.data_ptr()),
            ctypes.cast(ctypes.c_void_p(res.data_ptr()), _F64P), stream))
        return res


    def dipoles(self, positions, box_size, dipoles, transfer):
        """What ``cdf`` leaves in ``dipoles`` (D,4,3) and ``transfer_matrix`` (D,6,3,3) with ``dipole_flag = 1``
        (``hymd_bonded_dipoles``): contiguous device tensors of the position dtype, overwritten."""
        D = self.n_terms[2]
        if tuple(dipoles.shape) != (D, 4, 3) or tuple(transfer.shape) != (D, 6, 3, 3):
            raise ValueError(f"dipoles / transfer matrices must have shapes ({D}, 4, 3) / ({D}, 6, 3, 3)")
        for t in (dipoles, transfer):
            if t.dtype != positions.dtype or not t.is_contiguous() or not t.is_cuda:
                raise ValueError("dipoles / transfer matrices must be contiguous device tensors of the position dtype")
        box = (ctypes.c_double * 3)(*[float(b) for b in np.asarray(box_size).reshape(-1)[:3]])
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.hymd_bonded_dipoles(
            self._h, _lib.F64 if positions.dtype == torch.float64 else _lib.F32,
            ctypes.c_void_p(positions.data_ptr()), box, ctypes.c_void_p(dipoles.data_ptr()),
            ctypes.c_void_p(transfer.data_ptr()), stream))

    def redistribute(self, f_dipoles, transfer, f_beads):
        """``dipole_forces_redistribution`` (``hymd_dipole_redistribute``): ``f_beads`` (n,3) is overwritten."""
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.hymd_dipole_redistribute(
            self._h, _lib.F64 if f_beads.dtype == torch.float64 else _lib.F32,
            ctypes.c_void_p(f_dipoles.data_ptr()), ctypes.c_void_p(transfer.data_ptr()),
            ctypes.c_void_p(f_beads.data_ptr()), stream))

    def inner_step(self, x_in, x_out, vel, box_size, mass, kick_dt, n_kicks, drift_dt, force_out=None,
                   want_energies=True, cta=None):
        """One fused inner rRESPA step (``hymd_bonded_inner_step``): bonded forces at ``x_in``,
        ``n_kicks`` half kicks of ``vel`` (in place) and, if ``x_out`` is given (a different tensor),
        ``x_out = mod(x_in + drift_dt*vel, box)``.  ``force_out`` = optional 3-list of (n,3) tensors
        (bond, angle, dihedral; entries may be None).  Returns the (3,4) float64 device tensor
        {energy, pr_x, pr_y, pr_z} per kind (aliases an internal buffer) or None.  ``cta`` selects the
        evaluation strategy for this and later calls (see :meth:`set_cta`)."""
        if cta is not None and cta != self._cta:
            try:
                self.set_cta(cta)
            except _lib.HymdError:      # a CTA's term list does not fit in shared memory: per-particle
                self.set_cta(0)
                self._cta = cta         # do not retry every step
        n = self.n_particles
        tensors = [x_in, vel] + ([x_out] if x_out is not None else []) + \
            [f for f in (force_out or []) if f is not None]
        for t in tensors:
            if tuple(t.shape) != (n, 3) or t.dtype != x_in.dtype or not t.is_contiguous():
                raise ValueError(f"inner_step needs contiguous ({n}, 3) tensors of one dtype")
        if x_in.dtype not in (torch.float32, torch.float64):
            raise ValueError("positions must be float32 or float64")
        box = (ctypes.c_double * 3)(*[float(b) for b in np.asarray(box_size).reshape(-1)[:3]])
        fptr = None
        if force_out is not None:
            fptr = (ctypes.c_void_p * 3)(*[f.data_ptr() if f is not None else None for f in force_out])
        res = self._out12 if want_energies else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.hymd_bonded_inner_step(
            self._h, _lib.F64 if x_in.dtype == torch.float64 else _lib.F32,
            ctypes.c_void_p(x_in.data_ptr()), ctypes.c_void_p(x_out.data_ptr()) if x_out is not None else None,
            ctypes.c_void_p(vel.data_ptr()), box, float(mass), float(kick_dt), int(n_kicks), float(drift_dt),
            fptr, ctypes.cast(ctypes.c_void_p(res.data_ptr()), _F64P) if res is not None else None, stream))
        _lib.mark_written(vel)
        if x_out is not None:
            _lib.mark_written(x_out)
        return res


# topology cache: keyed by the identity and length of the caller's index arrays, so that the
# reference's call pattern (same prepare_bonds arrays every inner step) uploads once.  An entry dies
# with its arrays (weak references); at most _CACHE_MAX device topologies are kept alive.
_cache = {}
_CACHE_MAX = 6


def _topology(kind, n_particles, device, arrays, build):
    key = (kind, n_particles, str(device)) + tuple((id(a), len(a)) for a in arrays)
    hit = _cache.get(key)
    if hit is not None and all(r() is a for r, a in zip(hit[1], arrays)):
        return hit[0]
    topo = build()
    try:
        refs = [weakref.ref(a, lambda _r, k=key: _cache.pop(k, None)) for a in arrays]
    except TypeError:      # lists cannot be weak-referenced: no caching
        return topo
    while len(_cache) >= _CACHE_MAX:
        _cache.pop(next(iter(_cache)))
    _cache[key] = (topo, refs)
    return topo


def _device_io(r, f):
    """(device positions, device force buffer, copy-back callback or None)."""
    if isinstance(r, torch.Tensor) and r.is_cuda:
        pos = r.contiguous()
        if isinstance(f, torch.Tensor) and f.is_cuda and f.is_contiguous() and f.dtype == pos.dtype:
            return pos, f, None
        buf = torch.empty_like(pos)
        return pos, buf, (lambda: f.copy_(buf)) if isinstance(f, torch.Tensor) else (lambda: f.__setitem__(Ellipsis, buf.cpu().numpy()))
    if not torch.cuda.is_available():
        raise _lib.HymdError("hymd_b200.force needs a CUDA device (no CPU fallback)")
    arr = np.asarray(r)
    dt = torch.float64 if arr.dtype == np.float64 else torch.float32
    pos = torch.as_tensor(np.ascontiguousarray(arr), dtype=dt).cuda()
    buf = torch.empty_like(pos)

    def back():
        if isinstance(f, torch.Tensor):
            f.copy_(buf)
        else:
            f[...] = buf.cpu().numpy()
    return pos, buf, back


def _result(res, on_device, with_pr=True):
    if on_device:
        res = res.clone()      # the topology reuses its result buffer on the next call
        return (res[0], res[1:4]) if with_pr else res[0]
    host = res.cpu().numpy()
    return (float(host[0]), host[1:4].copy()) if with_pr else float(host[0])


def compute_bond_forces(f_bonds, r, box_size, a, b, r0, k):
    """``cbf``: harmonic two-particle bonds; writes ``f_bonds`` in place, returns
    ``(energy, bond_pr)``."""
    pos, buf, back = _device_io(r, f_bonds)
    topo = _topology(2, pos.shape[0], pos.device, (a, b, r0, k),
                     lambda: BondedTopology(pos.shape[0], bonds=(a, b, r0, k), device=pos.device))
    res = topo.forces(2, pos, box_size, buf)
    if back is not None:
        back()
    return _result(res, back is None)


def compute_angle_forces(f_angles, r, box_size, a, b, c, t0, k):
    """``caf``: harmonic three-particle angles; returns ``(energy, angle_pr)``."""
    pos, buf, back = _device_io(r, f_angles)
    topo = _topology(3, pos.shape[0], pos.device, (a, b, c, t0, k),
                     lambda: BondedTopology(pos.shape[0], angles=(a, b, c, t0, k), device=pos.device))
    res = topo.forces(3, pos, box_size, buf)
    if back is not None:
        back()
    return _result(res, back is None)


def _like_device(x, ref, shape):
    """Contiguous device tensor of ``ref``'s dtype for the caller's array ``x`` (a view when it already is one)."""
    if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == ref.dtype and x.is_contiguous() \
            and tuple(x.shape) == tuple(shape):
        return x, True
    return torch.empty(shape, dtype=ref.dtype, device=ref.device), False


def _store(dst, src):
    """``dst[...] = src`` for torch tensors and numpy arrays of any memory order."""
    if isinstance(dst, torch.Tensor):
        dst.copy_(src.reshape(dst.shape))
    else:
        dst[...] = src.cpu().numpy().reshape(dst.shape)


def compute_dihedral_forces(f_dihedrals, r, dipoles, transfer_matrix, box_size, a, b, c, d, coeff,
                            dtype, bb_index=None, dipole_flag=0):
    """``cdf``: ``dtype`` 0 (cosine series), 1 (combined bending-torsion: the series plus
    ``k(phi) (gamma - gamma_0(phi))^2`` on the angle a-b-c and, where ``bb_index`` = ``bonds_4_last`` is 1, on b-c-d)
    and 2 (improper); returns the energy.  ``dipoles`` (D,4,3) and ``transfer_matrix`` (D,6,3,3) are zeroed like
    the Fortran does (``compute_dihedral_forces.f90:27-28``) and, with ``dipole_flag = 1``, receive the
    reconstructed backbone dipoles and their transfer matrices (``dipole_reconstruction.f90:50-221``)."""
    pos, buf, back = _device_io(r, f_dihedrals)
    last = bb_index if bb_index is not None else np.zeros(len(a), dtype=np.int32)
    arrays = (a, b, c, d, coeff, dtype) + ((bb_index,) if bb_index is not None else ())
    topo = _topology(4, pos.shape[0], pos.device, arrays,
                     lambda: BondedTopology(pos.shape[0], dihedrals=(a, b, c, d, coeff, dtype, last),
                                            device=pos.device))
    res = topo.forces(4, pos, box_size, buf)
    D = topo.n_terms[2]
    if dipole_flag and topo.n_cbt and dipoles is not None and transfer_matrix is not None:
        dd, d_view = _like_device(dipoles, pos, (D, 4, 3))
        tt, t_view = _like_device(transfer_matrix, pos, (D, 6, 3, 3))
        topo.dipoles(pos, box_size, dd, tt)
        if not d_view:
            _store(dipoles, dd)
        if not t_view:
            _store(transfer_matrix, tt)
    else:
        for arr in (dipoles, transfer_matrix):
            if arr is None:
                continue
            if isinstance(arr, torch.Tensor):
                arr.zero_()
            else:
                arr[...] = 0
    if back is not None:
        back()
    return _result(res, back is None, with_pr=False)


def dipole_forces_redistribution(f_on_bead, f_dipoles, trans_matrices, a, b, c, d, type_array, last_bb,
                                 coeff=None):
    """``dipole_forces_redistribution`` (``hymd/force.py:855-880``): the electrostatic forces on the reconstructed
    dipole charges ``f_dipoles`` (D,4,3) carried to the backbone beads through the transfer matrices; ``f_on_bead``
    (N,3) is overwritten.  The topology is the one cached by ``compute_dihedral_forces`` for the same index arrays
    (pass ``coeff`` to hit that cache entry; otherwise a force-free topology is built once)."""
    n = f_on_bead.shape[0]
    dev_in = isinstance(f_on_bead, torch.Tensor) and f_on_bead.is_cuda
    device = f_on_bead.device if dev_in else torch.device(_default_device())
    D = len(a)
    if coeff is None:
        coeff = _zero_coeff(D)
    arrays = (a, b, c, d, coeff, type_array, last_bb)
    topo = _topology(4, n, device, arrays,
                     lambda: BondedTopology(n, dihedrals=(a, b, c, d, coeff, type_array, last_bb), device=device))
    if dev_in:
        ref = f_on_bead
    else:
        dt = torch.float64 if np.asarray(f_on_bead).dtype == np.float64 else torch.float32
        ref = torch.empty((n, 3), dtype=dt, device=device)

    def dev(x, shape):
        if isinstance(x, torch.Tensor):
            return x.to(device=device, dtype=ref.dtype).reshape(shape).contiguous()
        return torch.as_tensor(np.ascontiguousarray(np.asarray(x).reshape(shape)), dtype=ref.dtype).to(device)
    out = f_on_bead if dev_in and f_on_bead.is_contiguous() else torch.empty((n, 3), dtype=ref.dtype, device=device)
    topo.redistribute(dev(f_dipoles, (D, 4, 3)), dev(trans_matrices, (D, 6, 3, 3)), out)
    if out is not f_on_bead:
        _store(f_on_bead, out)


_zero_coeffs = {}


def _zero_coeff(D):
    if D not in _zero_coeffs:
        _zero_coeffs[D] = np.zeros((D, 6, 5))
    return _zero_coeffs[D]

"""Device versions of ``hymd/thermostat.py``.

=================================  ===============================
this module                        reference
=================================  ===============================
``csvr_thermostat``                ``thermostat.py:111-219``
``cancel_com_momentum``            ``thermostat.py:12-15``
``generate_initial_velocities``    ``thermostat.py:18-46``
``velocity_moments``               the ``comm.allreduce`` sums inside them
=================================  ===============================

Same names, argument order and in-place semantics.  The velocities are a torch CUDA tensor (updated in
place, no copies) or a numpy array (copied in and out).  The random numbers are drawn on the host
from the caller's ``prng`` in the reference's order (one Gaussian, one chi-squared per coupling
group), so a run seeded like the reference sees the same stochastic sequence; the kinetic energies,
the rescaling factor and the rescaling itself are evaluated on the device (``csrc/md.cu``) and the
host never waits for them: ``config.thermostat_work`` becomes a :class:`DeviceScalar` that is only
read back when converted with ``float()``.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

_F64P = ctypes.POINTER(ctypes.c_double)
_I32P = ctypes.POINTER(ctypes.c_int32)


class DeviceScalar:
    """A float64 scalar that lives on the GPU until somebody needs its value.

    ``config.thermostat_work`` is a plain float in the reference and its consumers treat it as one
    (``file_io.py:704`` stores it into an HDF5 dataset, ``file_io.py:751`` subtracts it from the total
    energy), so this class implements the numeric protocol: any arithmetic, comparison, formatting or
    ``numpy`` conversion reads the value back (one synchronisation, on print steps only) and yields a
    Python float."""

    __array_priority__ = 100

    def __init__(self, tensor):
        self.tensor = tensor

    def __float__(self):
        return float(self.tensor.item())

    def item(self):
        return float(self)

    def __repr__(self):
        return f"DeviceScalar({float(self)!r})"

    def __array__(self, dtype=None, copy=None):
        return np.asarray(float(self), dtype=dtype or np.float64)

    def __format__(self, spec):
        return format(float(self), spec)

    def __int__(self):
        return int(float(self))

    def __bool__(self):
        return bool(float(self))

    def __hash__(self):
        return hash(float(self))

    def __neg__(self):
        return -float(self)

    def __pos__(self):
        return float(self)

    def __abs__(self):
        return abs(float(self))

    def __round__(self, n=None):
        return round(float(self), n)


def _binary(name):
    import operator
    op = getattr(operator, name)

    def fwd(self, other):
        return op(float(self), float(other) if isinstance(other, DeviceScalar) else other)

    def rev(self, other):
        return op(float(other) if isinstance(other, DeviceScalar) else other, float(self))
    return fwd, rev


for _n in ("add", "sub", "mul", "truediv", "floordiv", "mod", "pow"):
    _f, _r = _binary(_n)
    setattr(DeviceScalar, f"__{_n}__", _f)
    setattr(DeviceScalar, f"__r{_n}__", _r)
for _n in ("lt", "le", "gt", "ge", "eq", "ne"):
    setattr(DeviceScalar, f"__{_n}__", _binary(_n)[0])
del _n, _f, _r


def _dtype_code(t):
    if t.dtype == torch.float64:
        return _lib.F64
    if t.dtype == torch.float32:
        return _lib.F32
    raise ValueError("velocities must be float32 or float64")


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _device_velocities(velocity):
    if isinstance(velocity, torch.Tensor) and velocity.is_cuda:
        if not velocity.is_contiguous():
            raise ValueError("velocities must be contiguous (N,3)")
        return velocity, None
    if not torch.cuda.is_available():
        raise _lib.HymdError("hymd_b200.thermostat needs a CUDA device (no CPU fallback)")
    arr = np.asarray(velocity)
    dt = torch.float64 if arr.dtype == np.float64 else torch.float32
    buf = torch.as_tensor(np.ascontiguousarray(arr), dtype=dt).cuda()

    def back():
        if isinstance(velocity, torch.Tensor):
            velocity.copy_(buf)
        else:
            velocity[...] = buf.cpu().numpy()
    return buf, back


_scratch = {}


def velocity_moments(vel, group=None, g=-1, allreduce=True):
    """(10,) float64 device tensor: {count, sum v (3), sum |v|^2} of the particles with
    ``group[i] == g`` (all if ``group`` is None) followed by the same for all particles; summed over
    the ranks of ``torch.distributed`` when initialized (the reference's ``comm.allreduce``)."""
    lib = _lib.load()
    key = str(vel.device)
    if key not in _scratch:
        _scratch[key] = torch.empty(int(lib.hymd_velocity_moments_scratch_doubles()), dtype=torch.float64,
                                    device=vel.device)
    out = torch.empty(10, dtype=torch.float64, device=vel.device)
    gp = ctypes.cast(ctypes.c_void_p(group.data_ptr()), _I32P) if group is not None else None
    _lib.check(lib.hymd_velocity_moments(
        _dtype_code(vel), ctypes.c_void_p(vel.data_ptr()), gp, int(g), int(vel.shape[0]),
        ctypes.cast(ctypes.c_void_p(_scratch[key].data_ptr()), _F64P),
        ctypes.cast(ctypes.c_void_p(out.data_ptr()), _F64P), _stream(vel.device)))
    if allreduce:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(out)
    return out


def kinetic_energy(vel, mass):
    """``0.5 * m * sum v^2`` over all ranks (``field.py:695``) as a :class:`DeviceScalar`."""
    return DeviceScalar(0.5 * mass * velocity_moments(vel)[9])


def _group_ids(names, config, device):
    """int32 device tensor: coupling group of every particle, -1 if none (``thermostat.py:177-183``).
    ``names`` is the reference's per-particle byte-string array, or an integer type-id array / tensor
    (then ``config.name_to_type_map`` translates the group names)."""
    groups = config.thermostat_coupling_groups
    if isinstance(names, torch.Tensor) or np.issubdtype(np.asarray(names).dtype, np.integer):
        types = names if isinstance(names, torch.Tensor) else torch.as_tensor(np.asarray(names))
        types = types.to(device).long()
        out = torch.full(types.shape, -1, dtype=torch.int32, device=device)
        for i, grp in enumerate(groups):
            for t in grp:
                out[types == int(config.name_to_type_map[t])] = i
        return out
    names = np.asarray(names)
    out = np.full(names.shape, -1, dtype=np.int32)
    for i, grp in enumerate(groups):
        for t in grp:
            out[names == np.bytes_(t)] = i
    return torch.as_tensor(out).to(device)


_group_cache = {}


def _random_gaussian(prng):
    return prng.normal()


def _random_chi_squared(prng, M):
    return prng.chisquare(M)


def csvr_thermostat(velocity, names, config, prng, comm=None, random_gaussian=_random_gaussian,
                    random_chi_squared=_random_chi_squared, remove_center_of_mass_momentum=True):
    """Canonical-sampling velocity rescaling, one coupling group after the other
    (``thermostat.py:177-219``).  Kept from the reference on purpose: with
    ``remove_center_of_mass_momentum=False`` or a single-particle group the kinetic energy of ALL
    particles enters and ALL velocities are rescaled (``thermostat.py:190-191, 216-217``)."""
    lib = _lib.load()
    vel, back = _device_velocities(velocity)
    if not any(config.thermostat_coupling_groups):
        config.thermostat_coupling_groups = [list(config.unique_names)]
    key = (id(names), str(vel.device), tuple(tuple(g) for g in config.thermostat_coupling_groups))
    hit = _group_cache.get(key)
    if hit is None or hit[0] is not names:
        if len(_group_cache) > 16:
            _group_cache.clear()
        grp = _group_ids(names, config, vel.device)
        counts = torch.bincount((grp + 1).long(), minlength=len(config.thermostat_coupling_groups) + 1)[1:]
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(counts)
        hit = (names, grp, [int(x) for x in counts.cpu()])
        _group_cache[key] = hit
    _, grp, counts = hit
    work = getattr(config, "thermostat_work", 0.0)
    if isinstance(work, DeviceScalar) and work.tensor.device == vel.device:
        work_t = work.tensor
    else:
        work_t = torch.full((1,), float(work or 0.0), dtype=torch.float64, device=vel.device)
    c = float(np.exp(-(config.time_step * config.respa_inner) / config.tau))
    kT15 = 1.5 * config.gas_constant * config.target_temperature
    n = int(vel.shape[0])
    for g in range(len(config.thermostat_coupling_groups)):
        mom = velocity_moments(vel, grp, g)
        R = float(random_gaussian(prng))
        SNf = float(random_chi_squared(prng, 3 * counts[g] - 1))
        _lib.check(lib.hymd_csvr_apply(
            _dtype_code(vel), ctypes.c_void_p(vel.data_ptr()),
            ctypes.cast(ctypes.c_void_p(grp.data_ptr()), _I32P), g, n,
            ctypes.cast(ctypes.c_void_p(mom.data_ptr()), _F64P), float(config.mass), kT15, c, R, SNf,
            1 if remove_center_of_mass_momentum else 0,
            ctypes.cast(ctypes.c_void_p(work_t.data_ptr()), _F64P), _stream(vel.device)))
    config.thermostat_work = DeviceScalar(work_t)
    if back is not None:
        back()
        config.thermostat_work = float(config.thermostat_work)
    return velocity


def cancel_com_momentum(velocities, config, comm=None):
    """``v -= sum(v) / n_particles`` (``thermostat.py:12-15``); in place, returns ``velocities``."""
    lib = _lib.load()
    vel, back = _device_velocities(velocities)
    mom = velocity_moments(vel)
    _lib.check(lib.hymd_cancel_com(_dtype_code(vel), ctypes.c_void_p(vel.data_ptr()), int(vel.shape[0]),
                                   ctypes.cast(ctypes.c_void_p(mom.data_ptr()), _F64P),
                                   float(config.n_particles), _stream(vel.device)))
    if back is not None:
        back()
    return velocities


def generate_initial_velocities(velocities, config, prng, comm=None):
    """``thermostat.py:18-46``: normal deviates with scale kT/m (drawn on the host from ``prng`` like
    the reference), centre-of-mass momentum removed, rescaled to the target kinetic energy."""
    kT_start = config.gas_constant * config.start_temperature
    n_local = int(velocities.shape[0])
    draw = prng.normal(loc=0, scale=kT_start / config.mass, size=(n_local, 3))
    if isinstance(velocities, torch.Tensor):
        velocities.copy_(torch.as_tensor(draw, dtype=velocities.dtype))
    else:
        velocities[...] = draw
    vel, back = _device_velocities(velocities)
    cancel_com_momentum(vel, config)
    K = 0.5 * config.mass * velocity_moments(vel)[9]
    factor = torch.sqrt(1.5 * config.n_particles * kT_start / K)
    vel.mul_(factor.to(vel.dtype))
    if back is not None:
        back()
    return velocities

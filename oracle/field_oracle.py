"""CPU restatement of HyMD's field-force cycle (``hymd/field.py``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Single rank: the
reference's results do not depend on the MPI decomposition beyond summation
order, so the oracle ignores layouts.

Every function follows the reference's sequence of operations (same number of
FFTs, same place where the filter is applied, real-space evaluation of v_ext)
so that it doubles as the timed "port" CPU baseline in ``bench.py``.
"""
from __future__ import annotations

import numpy as np

from . import pm_oracle as pmo


class FieldState:
    """What ``initialize_pm`` allocates (``field.py:48-86``), as plain numpy arrays."""

    def __init__(self, config, dtype=None):
        self.dtype = np.dtype(dtype if dtype is not None else (config.dtype or np.float64))
        self.mesh = pmo.mesh_tuple(config.mesh_size)
        self.box = np.asarray(config.box_size, dtype=np.float64)
        t = config.n_types
        self.phi = [None] * t                      # filtered densities after update_field
        self.phi_fourier = [None] * t              # H * r2c(phi)
        self.force_mesh = [[None] * 3 for _ in range(t)]
        self.v_ext = [None] * t
        self.phi_q = None
        self.phi_q_fourier = None
        self.psi = None
        self.elec_field = [None] * 3


def volume_per_cell(config):
    """``field.py:570-572``."""
    v = np.prod(np.asarray(config.box_size, dtype=np.float64))
    return float(v / np.prod(np.full(3, config.mesh_size)))


def update_field(state, hamiltonian, positions, types, config, m=None,
                 compute_potential=False, workers=1, use_c=True, mt=False):
    """``field.py:570-616``: densities -> filter -> v_ext -> filter -> -ik -> force meshes."""
    dt = state.dtype
    mesh, box = state.mesh, state.box
    dv = volume_per_cell(config)
    k = pmo.kgrid(mesh, box, dt)
    m = m if m is not None else (config.m or [1.0] * config.n_types)
    paint = _paint_mt if mt else pmo.cic_paint
    # loop A (field.py:573-578)
    for t in range(config.n_types):
        phi = paint(positions[types == t], m[t], mesh, box, dt, use_c)
        phi /= dt.type(dv)
        pf = pmo.r2c(phi, workers)
        pf = hamiltonian.H(k, pf).astype(pf.dtype, copy=False)
        state.phi_fourier[t] = pf
        state.phi[t] = pmo.c2r(pf, mesh, workers)
    # loop B (field.py:581-616)
    for t in range(config.n_types):
        v = np.asarray(hamiltonian.v_ext[t](state.phi), dtype=dt)
        vf = pmo.r2c(v, workers)
        vf = hamiltonian.H(k, vf).astype(vf.dtype, copy=False)
        for d in range(3):
            state.force_mesh[t][d] = pmo.c2r((-1j * k[d] * vf).astype(vf.dtype, copy=False),
                                             mesh, workers)
        if compute_potential:
            state.v_ext[t] = pmo.c2r(vf, mesh, workers)
    return state


def compute_field_force(state, positions, types, n_types, force=None, use_c=True, mt=False):
    """``field.py:197-200``: CIC interpolation of the force meshes, caller order."""
    if force is None:
        force = np.zeros((positions.shape[0], 3), dtype=state.dtype)
    readout = _readout_mt if mt else pmo.cic_readout
    for t in range(n_types):
        ind = types == t
        for d in range(3):
            force[ind, d] = readout(state.force_mesh[t][d], positions[ind], state.box, use_c)
    return force


def compute_self_energy_q(config, charges):
    """``field.py:231-238``."""
    conv = config.coulomb_constant / config.dielectric_const
    prefac = conv * np.sqrt(1.0 / (2.0 * np.pi * config.sigma * config.sigma))
    return float(prefac * np.sum(np.asarray(charges, dtype=np.float64) ** 2))


def update_field_force_q(state, hamiltonian, charges, positions, config, elec_forces=None,
                         workers=1, use_c=True, mt=False):
    """``field.py:356-403``: PME ("PIC_Spectral") potential, field and forces."""
    dt = state.dtype
    mesh, box = state.mesh, state.box
    dv = volume_per_cell(config)
    conv = config.coulomb_constant / config.dielectric_const
    k = pmo.kgrid(mesh, box, dt)
    k2 = pmo.knorm2_zeromode1(k)
    paint = _paint_mt if mt else pmo.cic_paint
    readout = _readout_mt if mt else pmo.cic_readout
    phi_q = paint(positions, np.asarray(charges, dtype=dt), mesh, box, dt, use_c)
    phi_q /= dt.type(dv)
    state.phi_q = phi_q
    pf = pmo.r2c(phi_q, workers)
    pf = hamiltonian.H(k, pf).astype(pf.dtype, copy=False)
    state.phi_q_fourier = pf
    psi_f = (4.0 * np.pi * conv * pf / k2).astype(pf.dtype, copy=False)
    state.psi = pmo.c2r(psi_f, mesh, workers)
    for d in range(3):
        ef = (-1j * k[d] * 4.0 * np.pi * conv * pf / k2).astype(pf.dtype, copy=False)
        state.elec_field[d] = pmo.c2r(ef, mesh, workers)
    if elec_forces is None:
        elec_forces = np.zeros((positions.shape[0], 3), dtype=dt)
    q = np.asarray(charges, dtype=dt)
    for d in range(3):
        elec_forces[:, d] = q * readout(state.elec_field[d], positions, box, use_c)
    return elec_forces


def compute_field_and_kinetic_energy(state, hamiltonian, velocity, config):
    """``field.py:688-703``: (field_energy, kinetic_energy, field_q_energy)."""
    dv = volume_per_cell(config)
    field_energy = pmo.csum(hamiltonian.w_0(state.phi) * dv)
    kinetic = 0.5 * config.mass * float(np.sum(np.asarray(velocity, dtype=np.float64) ** 2))
    if config.coulombtype == "PIC_Spectral":
        field_q = pmo.csum(hamiltonian.w_elec([state.phi_q, state.psi]) * dv)
    else:
        field_q = 0.0
    return field_energy, kinetic, field_q


def comp_laplacian(state, config, workers=1):
    """``field.py:406-425``: ``phi_laplacian[t][d] = c2r(-k_d^2 * phi_fourier[t])``."""
    k = pmo.kgrid(state.mesh, state.box, state.dtype)
    state.phi_laplacian = [[None] * 3 for _ in range(config.n_types)]
    for t in range(config.n_types):
        for d in range(3):
            tr = (-k[d] ** 2 * state.phi_fourier[t]).astype(state.phi_fourier[t].dtype, copy=False)
            state.phi_laplacian[t][d] = pmo.c2r(tr, state.mesh, workers)
    return state.phi_laplacian


def comp_pressure(state, hamiltonian, velocities, config, bond_pr=None, angle_pr=None, workers=1):
    """``pressure.py:84-200`` (single rank): the 18 pressure contributions
    ``[p_kin, p0, p1, p2x, p2y, p2z, bond(3), angle(3), dihedral(3), total(3)]``."""
    bond_pr = np.zeros(3) if bond_pr is None else np.asarray(bond_pr, dtype=np.float64)
    angle_pr = np.zeros(3) if angle_pr is None else np.asarray(angle_pr, dtype=np.float64)
    V = float(np.prod(np.asarray(config.box_size, dtype=np.float64)))
    dv = volume_per_cell(config)
    elec = config.coulombtype in ("PIC_Spectral", "PIC_Spectral_GPE") and state.psi is not None
    w = hamiltonian.w_0(state.phi) * dv                                   # pressure.py:89-95
    if elec:
        w = w + hamiltonian.w_elec([state.phi_q, state.psi]) * dv
    kinetic = 0.5 * config.mass * float(np.sum(np.asarray(velocities, dtype=np.float64) ** 2))
    p_kin = 2.0 / (3.0 * V) * kinetic                                     # pressure.py:98-99
    p0 = -1.0 / V * float(np.sum(w, dtype=np.float64))                    # pressure.py:102
    if elec:                                                              # pressure.py:105-108
        v_bar = np.array([hamiltonian.V_bar[t]([state.phi, state.psi]) for t in range(config.n_types)])
    else:
        v_bar = np.array([hamiltonian.V_bar_0[t](state.phi) for t in range(config.n_types)])
    p1 = float(np.sum((dv / V) * v_bar * np.asarray(state.phi), dtype=np.float64))   # pressure.py:110
    comp_laplacian(state, config, workers)                                # pressure.py:113-119
    lap = np.asarray(state.phi_laplacian)                                 # (T, 3, Nx, Ny, Nz)
    p2 = np.sum(dv / V * config.sigma ** 2 * np.repeat(v_bar[:, np.newaxis], 3, axis=1) * lap,
                axis=(0, 2, 3, 4), dtype=np.float64)                      # pressure.py:121-127
    p_bond, p_angle, p_dih = bond_pr / V, angle_pr / V, np.zeros(3)      # pressure.py:130-136
    p_tot = p_kin + p0 + p1 + p2 + p_bond + p_angle + p_dih              # pressure.py:154
    return np.array([p_kin, p0, p1, p2[0], p2[1], p2[2], *p_bond, *p_angle, *p_dih, *p_tot])


# --- multi-threaded CIC (CPU baseline only; same arithmetic) ----------------------------

def _paint_mt(pos, mass, mesh, box, dtype, use_c=True):
    import ctypes
    lib = pmo._load_clib()
    if lib is None:
        return pmo.cic_paint(pos, mass, mesh, box, dtype, use_c)
    dtype = np.dtype(dtype)
    nx, ny, nz = pmo.mesh_tuple(mesh)
    pos = np.ascontiguousarray(pos, dtype=dtype).reshape(-1, 3)
    n = pos.shape[0]
    m = np.ascontiguousarray(np.broadcast_to(np.asarray(mass, dtype=dtype), (n,)))
    out = np.zeros((nx, ny, nz), dtype=dtype)
    fn = lib.cic_paint_mt_f64 if dtype == np.float64 else lib.cic_paint_mt_f32
    b = np.asarray(box, dtype=np.float64)
    fn(pos.ctypes.data_as(ctypes.c_void_p), m.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(n),
       ctypes.c_int(nx), ctypes.c_int(ny), ctypes.c_int(nz), ctypes.c_double(b[0]),
       ctypes.c_double(b[1]), ctypes.c_double(b[2]), out.ctypes.data_as(ctypes.c_void_p))
    return out


def _readout_mt(field, pos, box, use_c=True):
    import ctypes
    lib = pmo._load_clib()
    if lib is None or not field.flags.c_contiguous:
        return pmo.cic_readout(field, pos, box, use_c)
    dtype = field.dtype
    pos = np.ascontiguousarray(pos, dtype=dtype).reshape(-1, 3)
    n = pos.shape[0]
    out = np.empty(n, dtype=dtype)
    fn = lib.cic_readout_mt_f64 if dtype == np.float64 else lib.cic_readout_mt_f32
    b = np.asarray(box, dtype=np.float64)
    fn(field.ctypes.data_as(ctypes.c_void_p), pos.ctypes.data_as(ctypes.c_void_p),
       ctypes.c_long(n), ctypes.c_int(field.shape[0]), ctypes.c_int(field.shape[1]),
       ctypes.c_int(field.shape[2]), ctypes.c_double(b[0]), ctypes.c_double(b[1]),
       ctypes.c_double(b[2]), out.ctypes.data_as(ctypes.c_void_p))
    return out

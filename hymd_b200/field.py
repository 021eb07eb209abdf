"""Drop-in replacement for the hot path of ``hymd/field.py`` on one or more B200 GPUs.

Same function names, positional argument order and in-place output conventions as the
reference, so ``main.py`` / ``integrator.py`` can call these unchanged:

====================================  =========================
this module                           reference
====================================  =========================
``initialize_pm``                     ``field.py:10-149``
``update_field``                      ``field.py:428-616``
``compute_field_force``               ``field.py:152-200``
``update_field_force_q``              ``field.py:241-403``
``compute_self_energy_q``             ``field.py:203-238``
``compute_field_and_kinetic_energy``  ``field.py:619-703``
``comp_laplacian``                    ``field.py:406-425``
``domain_decomposition``              ``field.py:1115-1178``
====================================  =========================

Particle arrays (``positions (N,3)``, ``types (N,)``, ``charges (N,)``, ``force (N,3)``) may be
torch CUDA tensors (fast path, no copies) or numpy arrays in C or Fortran order (compatibility
path: copied to the device and, for outputs, back).  All arithmetic happens in
``libhymd_b200.so``; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .hamiltonian import energy_parameters
from .pm import MeshField, ParticleMesh, UnusedField


def initialize_pm(pmesh, config, comm=None):
    """Create the particle-mesh context and the field handles (``field.py:10-149``).

    ``pmesh`` (the pmesh module in the reference) is accepted for signature compatibility and
    ignored.  Returns ``(pm, field_list, elec_common_list, coulomb_list)`` with the reference's
    list layouts (``field.py:139-147, 66-75, 83-86``)."""
    dtype = "f8" if np.dtype(config.dtype) == np.float64 else "f4"
    coulombtype = getattr(config, "coulombtype", None)
    pm = ParticleMesh(config.mesh_size, BoxSize=config.box_size, dtype=dtype, comm=comm,
                      config=config)
    T = config.n_types
    phi = [pm.field(_lib.FIELD_PHI, t) for t in range(T)]
    phi_fourier = [pm.field(_lib.FIELD_PHI_FOURIER, t, kind="complex") for t in range(T)]
    force_on_grid = [[pm.field(_lib.FIELD_FORCE_MESH, t, d) for d in range(3)] for t in range(T)]
    v_ext_fourier = [UnusedField(f"v_ext_fourier[{i}]") for i in range(4)]
    v_ext = [pm.field(_lib.FIELD_V_EXT, t) for t in range(T)]
    phi_transfer = [UnusedField(f"phi_transfer[{i}]") for i in range(3)]
    phi_laplacian = [[pm.field(_lib.FIELD_PHI_LAPLACIAN, t, d) for d in range(3)] for t in range(T)]
    field_list = [phi, phi_fourier, force_on_grid, v_ext_fourier, v_ext, phi_transfer,
                  phi_laplacian]
    elec_common_list = [None, None, None, None]
    coulomb_list = []
    if coulombtype == "PIC_Spectral":
        phi_q = pm.field(_lib.FIELD_PHI_Q)
        phi_q_fourier = pm.field(_lib.FIELD_PHI_Q_FOURIER, kind="complex")
        psi = pm.field(_lib.FIELD_PSI)
        elec_field = [pm.field(_lib.FIELD_ELEC_FIELD, 0, d) for d in range(3)]
        elec_common_list = [phi_q, phi_q_fourier, psi, elec_field]
        coulomb_list = [[UnusedField(f"elec_field_fourier[{d}]") for d in range(3)],
                        UnusedField("psi_fourier")]
    if coulombtype == "PIC_Spectral_GPE":          # list layouts of field.py:66-75, 88-137
        phi_q = pm.field(_lib.FIELD_PHI_Q)
        psi = pm.field(_lib.FIELD_PSI)
        elec_common_list = [phi_q, UnusedField("phi_q_fourier"), psi,
                            [UnusedField(f"elec_field[{d}]") for d in range(3)]]
        coulomb_list = [
            pm.field(_lib.FIELD_GPE_EPS), UnusedField("phi_eps_fourier"),
            [UnusedField(f"phi_eta[{d}]") for d in range(3)],
            [UnusedField(f"phi_eta_fourier[{d}]") for d in range(3)],
            UnusedField("phi_pol"), UnusedField("phi_pol_prev"), pm.field(_lib.FIELD_GPE_ELEC_DOT),
            UnusedField("elec_field_contrib"), [pm.field(_lib.FIELD_GPE_VBAR, t) for t in range(T)],
            [UnusedField(f"Vbar_elec_fourier[{t}]") for t in range(T)],
            [[UnusedField(f"force_mesh_elec[{t}][{d}]") for d in range(3)] for t in range(T)],
            [[UnusedField(f"force_mesh_elec_fourier[{t}][{d}]") for d in range(3)] for t in range(T)],
        ]
    return (pm, field_list, elec_common_list, coulomb_list)


def update_field(phi, phi_laplacian, phi_transfer, layouts, force_mesh, hamiltonian, pm,
                 positions, types, config, v_ext, phi_fourier, v_ext_fourier, m,
                 compute_potential=False):
    """Densities -> filter -> external potential -> force meshes (``field.py:428-616``).

    One counting sort of all particles, one deterministic CIC paint of every type, T forward
    FFTs, one fused k-space kernel and 3U inverse FFTs (U = distinct rows of the interaction
    matrix).  With ``compute_potential`` the filtered densities ``phi`` and the potentials
    ``v_ext`` are materialized as well (``field.py:578, 615-616``); otherwise they are computed
    lazily when a handle's ``.value`` is read."""
    pm.sync_interaction(hamiltonian, config, m)
    pm.sort(positions, types, force=True, cycle=1 if compute_potential else 0)


def _output_buffer(pm, out, n):
    """Device tensor the kernel writes into, and a callback copying it back if needed."""
    if isinstance(out, torch.Tensor) and out.device == pm.device and out.dtype == pm.dtype \
            and out.is_contiguous() and tuple(out.shape) == (n, 3):
        return out, None
    buf = torch.empty((n, 3), dtype=pm.dtype, device=pm.device)
    if isinstance(out, torch.Tensor):
        return buf, lambda: out.copy_(buf)

    def back():
        pm.to_host(buf, out, "forces")
    return buf, back


def compute_field_force(layouts, r, force_mesh, force, types, n_types):
    """CIC interpolation of the force meshes at the particle positions, written in place into
    ``force`` in the caller's particle order (``field.py:152-200``)."""
    pm = force_mesh[0][0].pm
    pm.sort(r, types)
    n = pm._n_local
    buf, back = _output_buffer(pm, force, n)
    _lib.check(pm.lib.hymd_readout(pm._ctx, ctypes.c_void_p(buf.data_ptr()), pm.stream))
    if back is not None:
        back()


def compute_self_energy_q(config, charges, comm=None):
    """Ewald self energy (``field.py:203-238``)."""
    conv = config.coulomb_constant / config.dielectric_const
    prefac = conv * np.sqrt(1.0 / (2.0 * np.pi * config.sigma * config.sigma))
    if isinstance(charges, torch.Tensor):
        s = (charges.double() ** 2).sum()
    else:
        s = torch.tensor(float(np.sum(np.asarray(charges, dtype=np.float64) ** 2)),
                         dtype=torch.float64)
    return float(prefac * _allreduce(s))


def update_field_force_q(charges, phi_q, phi_q_fourier, psi, psi_fourier, elec_field_fourier,
                         elec_field, elec_forces, layout_q, hamiltonian, pm, positions, config):
    """PME electrostatics (``field.py:241-403``): charge density, Poisson solve with the
    Gaussian filter, E = -grad psi, forces q*E written in place into ``elec_forces``.

    Called with mesh arguments that are not the handles of ``initialize_pm`` -- the peptide-dipole call of
    ``main.py:1060-1095`` passes ``pm.create`` meshes and the reconstructed dipole charges / positions -- the
    cycle runs in a second context (``pm.secondary_pme``), so the charge density and potential of the real
    charges, which the energy print reads afterwards, are left alone.  The passed meshes are opaque to
    ``main.py`` and are not filled."""
    if not isinstance(phi_q, MeshField) and getattr(pm, "pme", False) and not getattr(pm, "gpe", False):
        sec = pm.secondary_pme(config)
        sec.sort(positions, None, charges)
        n = sec._n_local
        buf, back = _output_buffer(sec, elec_forces, n)
        _lib.check(sec.lib.hymd_pme_cycle(sec._ctx, ctypes.c_void_p(buf.data_ptr()), 0, sec.stream))
        if back is not None:
            back()
        return
    pm.sync_interaction(hamiltonian, config)
    pm.sort(positions, None, charges)
    n = pm._n_local
    buf, back = _output_buffer(pm, elec_forces, n)
    _lib.check(pm.lib.hymd_pme_cycle(pm._ctx, ctypes.c_void_p(buf.data_ptr()), 0, pm.stream))
    if back is not None:
        back()


_CONVERGENCE = {None: 0, "max_diff": 0, "csum": 1, "euclidean_norm": 2}


def update_field_force_q_GPE(conv_fun, phi, types, charges, phi_q, phi_q_fourier, phi_eps, phi_eps_fourier,
                             phi_eta, phi_eta_fourier, phi_pol_prev, phi_pol, elec_field, elec_forces,
                             elec_field_contrib, psi, Vbar_elec, Vbar_elec_fourier, force_mesh_elec,
                             force_mesh_elec_fourier, hamiltonian, layout_q, layouts, pm, positions, config,
                             comm=None):
    """General-Poisson-equation electrostatics (``field.py:964-1112``): dielectric field from the type
    densities of the last ``update_field``, polarisation-charge iteration, potential, field, per-type
    electrostatic potential and the forces, written in place into ``elec_forces``.  ``conv_fun`` (the
    reference's closure over ``config.convergence_type``, ``main.py:141-163``) is accepted and ignored:
    the convergence measure is evaluated on the device according to ``config.convergence_type``.
    Returns ``(Vbar_elec, phi_eps, elec_dot)`` like the reference (``elec_dot`` as a mesh handle)."""
    pm.sync_interaction(hamiltonian, config)
    pm.sort(positions, None, charges)
    n = pm._n_local
    prm = _lib.HymdGpeParams()
    prm.struct_size = ctypes.sizeof(_lib.HymdGpeParams)
    prm.convergence_type = _CONVERGENCE[getattr(config, "convergence_type", None)]
    prm.max_iter = 100
    prm.pol_mixing = float(config.pol_mixing if getattr(config, "pol_mixing", None) is not None else 0.6)
    prm.conv_crit = float(config.conv_crit if getattr(config, "conv_crit", None) is not None else 1e-6)
    prm.coulomb_constant = float(config.coulomb_constant)
    for t in range(config.n_types):
        prm.dielectric_type[t] = float(config.dielectric_type[t])
        prm.type_charges[t] = float(config.type_charges[t])
    buf, back = _output_buffer(pm, elec_forces, n)
    iters = ctypes.c_int32(0)
    _lib.check(pm.lib.hymd_gpe_cycle(pm._ctx, ctypes.byref(prm), ctypes.c_void_p(buf.data_ptr()),
                                     ctypes.byref(iters), pm.stream))
    pm.gpe_iterations = int(iters.value)
    if back is not None:
        back()
    return Vbar_elec, phi_eps, pm.field(_lib.FIELD_GPE_ELEC_DOT)


def compute_field_energy_q_GPE(config, phi_eps, field_q_energy, dot_elec, comm=None):
    """``dV * eps_0 / 2 * sum(phi_eps * |E|^2)`` summed over ranks (``field.py:706-760``)."""
    pm = phi_eps.pm
    out = ctypes.c_double(0.0)
    _lib.check(pm.lib.hymd_gpe_energy(pm._ctx, float(config.coulomb_constant), ctypes.byref(out), pm.stream))
    return float(_allreduce(torch.tensor([out.value], dtype=torch.float64))[0])


def compute_field_and_kinetic_energy(phi, phi_q, psi, velocity, hamiltonian, positions, types,
                                     v_ext, config, layouts, comm=None):
    """``(field_energy, kinetic_energy, field_q_energy)`` (``field.py:619-703``)."""
    pm = phi[0].pm
    pme = getattr(config, "coulombtype", None) == "PIC_Spectral" and pm._sorted_with_charges
    pm._materialize(want_phi=True, want_psi=pme)
    params = energy_parameters(hamiltonian)
    if params is not None:
        chi, kappa, rho0, a = params
        out = (ctypes.c_double * 2)()
        chi = np.ascontiguousarray(chi, dtype=np.float64)
        _lib.check(pm.lib.hymd_field_energy(
            pm._ctx, chi.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), kappa, rho0, a, out,
            pm.stream))
        loc = torch.tensor([out[0], out[1]], dtype=torch.float64)
    else:  # user-defined affine functional: evaluate its own w_0 on the device tensors
        dv = float(np.prod(pm.BoxSize) / np.prod(pm.Nmesh))
        w = hamiltonian.w_0([f.value.double() for f in phi]).sum() * dv
        wq = (0.5 * phi_q.value.double() * psi.value.double()).sum() * dv if pme else torch.zeros(())
        loc = torch.stack([w.cpu(), wq.cpu().double()])
    if isinstance(velocity, torch.Tensor):
        kin = 0.5 * config.mass * (velocity.double() ** 2).sum().cpu()
    else:
        kin = torch.tensor(0.5 * config.mass * float(np.sum(np.asarray(velocity, dtype=np.float64) ** 2)),
                           dtype=torch.float64)
    tot = _allreduce(torch.stack([loc[0], loc[1], kin.reshape(())]))
    field_energy, field_q, kinetic = float(tot[0]), float(tot[1]), float(tot[2])
    if pme:
        field_q_energy = field_q - float(getattr(config, "self_energy", 0.0) or 0.0)
    else:
        field_q_energy = 0.0
    return field_energy, kinetic, field_q_energy


def comp_laplacian(phi_fourier, phi_transfer, phi_laplacian, hamiltonian, config):
    """``phi_laplacian[t][d] = c2r(-k_d^2 * phi_fourier[t])`` (``field.py:406-425``): one k-space
    kernel for all types and directions and 3T inverse transforms, from the density spectra of the
    last ``update_field``.  ``phi_transfer`` (the reference's scratch) is not used."""
    phi_laplacian[0][0].pm.laplacian()


def _first_atom_positions(pm, positions, molecules):
    """Routing positions with molecules: every atom follows the first atom of its molecule
    (``field.py:1156-1163``)."""
    mol = torch.as_tensor(np.asarray(molecules)) if not isinstance(molecules, torch.Tensor) \
        else molecules
    mol = mol.to(pm.device).long().reshape(-1)
    pos = pm.as_device(positions)
    n = mol.shape[0]
    uniq, inv = torch.unique(mol, return_inverse=True)
    first = torch.full((uniq.shape[0],), n, dtype=torch.long, device=pm.device)
    first.scatter_reduce_(0, inv, torch.arange(n, device=pm.device), reduce="amin")
    return pos[first[inv]]


def _cell_order(pm, route, molecules=None, block=0):
    """Stable permutation sorting the local particles by the mesh cell of their routing
    position (x slowest, z fastest = the storage order of the meshes).

    ``block > 0`` (``HYMD_B200_DD_BLOCK``, opt-in, needs ``molecules``): the primary key becomes (block
    of ``block``^3 cells, single-bead molecule?), the cell order is kept inside every such group.  The
    CTAs of the bonded kernels then hold either chain beads or solvent, not a mix,
    while the field kernels keep their locality at block granularity."""
    mesh = torch.as_tensor(np.asarray(pm.Nmesh, dtype=np.int64), device=pm.device)
    scale = torch.as_tensor(np.asarray(pm.Nmesh, dtype=np.float64) / np.asarray(pm.BoxSize, dtype=np.float64),
                            device=pm.device)
    c = torch.remainder(torch.floor(route.double() * scale).long(), mesh)
    key = (c[:, 0] * mesh[1] + c[:, 1]) * mesh[2] + c[:, 2]
    if block > 0 and molecules is not None:
        mol = torch.as_tensor(np.asarray(molecules)) if not isinstance(molecules, torch.Tensor) else molecules
        mol = mol.to(pm.device).long().reshape(-1)
        _, inv, counts = torch.unique(mol, return_inverse=True, return_counts=True)
        single = (counts[inv] == 1).long()
        nb = (mesh + block - 1) // block
        blk = ((c[:, 0] // block) * nb[1] + c[:, 1] // block) * nb[2] + c[:, 2] // block
        key = (blk * 2 + single) * (mesh[0] * mesh[1] * mesh[2]) + key
    return torch.sort(key, stable=True).indices


def _take_rows(a, perm):
    if isinstance(a, torch.Tensor):
        return a[perm.to(a.device)]
    return np.asarray(a)[perm.cpu().numpy()]


def domain_decomposition(positions, pm, *args, molecules=None, bonds=None, topol=False, verbose=0,
                         comm=None):
    """Re-home particles on the rank owning their slab (``field.py:1115-1178``) and hand the
    per-particle arrays back in mesh-cell order.

    Like the reference's ``Layout.exchange`` the call returns every array permuted identically;
    the order itself is implementation defined there (it depends on the rank layout).  Here it
    is the cell order of the routing position (molecules stay contiguous and keep their internal
    order: all atoms are keyed by the first atom, stable sort), so that until the next call the
    caller's particle order stays close to the order the field kernels stream in: the position
    gather of the cell binning and the force write-back of ``compute_field_force`` then touch
    neighbouring addresses.  ``HYMD_B200_DD_CELL_ORDER=0`` keeps the incoming order instead."""
    import os
    if molecules is not None:
        if not topol:
            assert bonds is not None, "bonds must be provided with molecules"
            args = (*args, bonds, molecules)
        else:
            args = (*args, molecules)
    arrays = (positions, *args)
    if pm.world_size > 1:
        route = None if molecules is None else _first_atom_positions(pm, positions, molecules)
        arrays = pm.migrate(positions, *args, routing_positions=route)
    if os.environ.get("HYMD_B200_DD_CELL_ORDER", "1") != "0" and len(arrays[0]) > 0:
        if molecules is None:
            route = pm.as_device(arrays[0])
        else:
            route = _first_atom_positions(pm, arrays[0], arrays[-1])
        perm = _cell_order(pm, route, None if molecules is None else arrays[-1],
                           int(os.environ.get("HYMD_B200_DD_BLOCK", "0")))
        arrays = tuple(_take_rows(a, perm) for a in arrays)
        pm.reset_order()
    return tuple(arrays)


def _allreduce(t):
    from . import _world
    return _world.current().allreduce(t).cpu()

"""``comp_pressure`` of ``hymd/pressure.py`` (84-200) on the device-resident fields.

Same signature and the same 18-entry result as the reference::

    [p_kin, p0, p1, p2x, p2y, p2z, bond x/y/z, angle x/y/z, dihedral x/y/z (zeros), total x/y/z]

The field terms are reductions over mesh cells done by ``libhymd_b200.so`` in double precision
with a fixed summation order (``hymd_field_energy``, ``hymd_field_pressure``); the Laplacians
come from ``hymd_b200.field.comp_laplacian``.  Like the reference (``pressure.py:198``) the
per-rank vector is summed over ranks at the end.
"""
from __future__ import annotations

import numpy as np
import torch

from .field import _allreduce, comp_laplacian, compute_field_and_kinetic_energy
from .hamiltonian import affine_parameters


def _host3(x):
    """bond_pr / angle_pr as 3 host doubles (hymd_b200.force returns them as device tensors when the
    positions live on the GPU)."""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float64).reshape(3)


def comp_pressure(phi, phi_q, psi, hamiltonian, velocities, config, phi_fourier, phi_laplacian,
                  phi_transfer, positions, bond_pr, angle_pr, comm=None):
    pm = phi[0].pm
    T = config.n_types
    V = float(np.prod(np.asarray(config.box_size, dtype=np.float64)))
    dv = V / float(np.prod(np.full(3, config.mesh_size)))
    elec = getattr(config, "coulombtype", None) in ("PIC_Spectral",) and psi is not None \
        and pm._sorted_with_charges
    # kinetic term and term 1 (pressure.py:89-102): -(1/V) sum_cells (w_0 + w_elec) dV is minus the
    # field energies of compute_field_and_kinetic_energy (which are already summed over ranks)
    e_field, e_kin, e_q = compute_field_and_kinetic_energy(
        phi, phi_q, psi, velocities, hamiltonian, positions, None, None, config, None, comm)
    p_kin = 2.0 / (3.0 * V) * e_kin
    p0 = -(e_field + (e_q if elec else 0.0)) / V
    # terms 2 and 3 (pressure.py:105-127): V_bar is the affine potential of the functional
    # (hamiltonian.py:157-186, 271-301, 423-475), V_bar_t = c_t + sum_j A_tj phi_j (+ q_t psi)
    comp_laplacian(phi_fourier, phi_transfer, phi_laplacian, hamiltonian, config)
    A, c = affine_parameters(hamiltonian, T)
    q = None
    if elec:
        q = np.asarray(list(config.type_charges), dtype=np.float64)
    sums = torch.as_tensor(pm.field_pressure_sums(A, c, q), dtype=torch.float64)
    sums = _allreduce(sums).numpy()
    p1 = dv / V * sums[0]
    p2 = dv / V * float(config.sigma) ** 2 * sums[1:4]
    # the reference adds the per-rank bonded terms and sums the whole vector over ranks; the field
    # and kinetic terms above are already global, so only the bonded inputs are reduced here
    bonded = np.concatenate([_host3(bond_pr), _host3(angle_pr)])
    if pm.world_size > 1:
        bonded = _allreduce(torch.as_tensor(bonded)).numpy()
    p_bond, p_angle = bonded[:3] / V, bonded[3:] / V
    p_dihedral = np.zeros(3)
    p_tot = p_kin + p0 + p1 + p2 + p_bond + p_angle + p_dihedral
    return np.array([p_kin, p0, p1, p2[0], p2[1], p2[2], *p_bond, *p_angle, *p_dihedral, *p_tot])

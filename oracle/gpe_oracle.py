"""CPU restatement of HyMD's general-Poisson-equation electrostatics (``coulombtype="PIC_Spectral_GPE"``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Groundwork for SURVEY.md section 8 row f3: the
product does not implement this branch yet (``hymd_b200.field.initialize_pm`` raises); the oracle and
its golden vectors are what the device path will be checked against.

Follows ``hymd/field.py:964-1112`` (``update_field_force_q_GPE``: dielectric field from the type
densities, polarisation-charge fixed-point iteration, potential, field, per-type electrostatic
external potential and forces) and ``hymd/field.py:745-760`` (``compute_field_energy_q_GPE``) on a
single rank.  Pinned against the reference's own function executed over ``oracle/pmesh_standin.py``
(``tests/golden/make_reference_golden.py`` -> ``tests/golden/gpe_golden.npz``,
``tests/test_oracle_gpe.py``).

Kept from the reference on purpose: the masks ``where=np.abs(x > 1e-6)`` (a boolean, i.e. ``x > 1e-6``:
cells at or below the threshold keep the previous content of the output array); ``phi_q`` is divided by
``phi_eps`` in place before the iteration; the iteration starts from the caller's ``phi_pol_prev``,
which the reference never writes back (zeros at every call in ``main.py``).
"""
from __future__ import annotations

import numpy as np

from . import pm_oracle as pmo


class GpeState:
    """The GPE meshes ``initialize_pm`` allocates (``field.py:88-137``), zero-initialised."""

    def __init__(self, mesh, n_types, dtype=np.float64):
        z = lambda: np.zeros(pmo.mesh_tuple(mesh), dtype=dtype)  # noqa: E731
        self.phi_q, self.phi_eps, self.phi_pol_prev, self.psi = z(), z(), z(), z()
        self.phi_eta = [z(), z(), z()]
        self.elec_field = [z(), z(), z()]
        self.elec_field_contrib = z()
        self.Vbar_elec = [z() for _ in range(n_types)]
        self.force_mesh_elec = [[z() for _ in range(3)] for _ in range(n_types)]
        self.elec_dot = z()
        self.iterations = 0


def update_field_force_q_GPE(st, phi, types, charges, positions, hamiltonian, config, conv="max_diff"):
    """``field.py:1003-1112``.  ``phi`` = filtered type densities of the last ``update_field``;
    returns the (N,3) electrostatic forces; ``st`` holds every mesh the reference updates."""
    mesh, box = st.phi_q.shape, np.asarray(config.box_size, dtype=np.float64)
    dt = st.phi_q.dtype
    dv = float(np.prod(box) / np.prod(mesh))
    k = pmo.kgrid(mesh, box, dt)
    k2 = pmo.knorm2_zeromode1(k)
    T = config.n_types
    eps_t = np.asarray(config.dielectric_type, dtype=np.float64)
    # smeared charge density (field.py:1006-1010)
    phi_q = pmo.cic_paint(positions, np.asarray(charges, dtype=dt), mesh, box, dt, use_c=False) / dt.type(dv)
    phi_q = pmo.c2r(hamiltonian.H(k, pmo.r2c(phi_q)), mesh)
    # dielectric field (field.py:1012-1020)
    num = np.zeros(mesh, dtype=dt)
    den = np.zeros(mesh, dtype=dt)
    for t in range(T):
        num = num + eps_t[t] * phi[t]
        den = den + phi[t]
    np.divide(num, den, where=den > 1e-6, out=st.phi_eps)
    phi_eps_f = pmo.r2c(st.phi_eps)
    np.divide(phi_q, st.phi_eps, where=st.phi_eps > 1e-6, out=phi_q)
    st.phi_q = phi_q
    for d in range(3):                                               # field.py:1024-1035
        eta = pmo.c2r(1j * k[d] * phi_eps_f, mesh)
        np.divide(eta, st.phi_eps, where=st.phi_eps > 1e-6, out=eta)
        st.phi_eta[d] = eta
    # polarisation-charge iteration (field.py:1037-1064)
    w = config.pol_mixing
    i, delta = 0, 1.0
    pol_prev = st.phi_pol_prev
    pol = pol_prev
    while i < 100 and delta > config.conv_crit:
        f = pmo.r2c(phi_q + pol_prev)
        for d in range(3):
            st.elec_field[d] = pmo.c2r(f * (-1j * k[d]) / k2, mesh)
        pol = -(st.phi_eta[0] * st.elec_field[0] + st.phi_eta[1] * st.elec_field[1]
                + st.phi_eta[2] * st.elec_field[2])
        pol = w * pol + (1.0 - w) * pol_prev
        diff = np.abs(pol - pol_prev)
        if conv == "max_diff":
            delta = float(np.max(diff))
        elif conv == "csum":
            delta = float(np.sum(diff))
        else:
            delta = float(np.sum(np.abs(diff) ** 2))
        pol_prev = pol.copy()
        i += 1
    st.iterations = i      # the caller's phi_pol_prev mesh is NOT updated: the reference rebinds its local
                           # name (field.py:1062), so every call starts the iteration from the same field
    # potential and field (field.py:1066-1084)
    eps0_inv = config.coulomb_constant * 4 * np.pi
    f = pmo.r2c(eps0_inv * (phi_q + pol)) / k2
    st.psi = pmo.c2r(f, mesh)
    for d in range(3):
        st.elec_field[d] = pmo.c2r(-1j * k[d] * f, mesh)
    st.elec_dot = st.elec_field[0] ** 2 + st.elec_field[1] ** 2 + st.elec_field[2] ** 2
    np.divide(st.elec_dot, den, where=den > 1e-6, out=st.elec_field_contrib)
    # per-type electrostatic external potential and forces (field.py:1086-1111)
    forces = np.zeros((len(positions), 3), dtype=dt)
    for t in range(T):
        st.Vbar_elec[t] = (config.type_charges[t] * st.psi
                           - (0.5 / eps0_inv) * (eps_t[t] - st.phi_eps) * st.elec_field_contrib)
    for t in range(T):
        vf = hamiltonian.H(k, pmo.r2c(st.Vbar_elec[t]))
        ind = types == t
        for d in range(3):
            st.force_mesh_elec[t][d] = pmo.c2r(-1j * k[d] * vf, mesh)
            forces[ind, d] = pmo.cic_readout(st.force_mesh_elec[t][d], positions[ind], box, use_c=False)
    return forces


def compute_field_energy_q_GPE(st, config):
    """``field.py:753-760``: ``dV * eps_0/2 * sum(phi_eps * |E|^2)``."""
    box = np.asarray(config.box_size, dtype=np.float64)
    dv = float(np.prod(box) / np.prod(st.phi_q.shape))
    eps_0 = 1.0 / (config.coulomb_constant * 4 * np.pi)
    return dv * (0.5 * eps_0) * float(np.sum(st.phi_eps * st.elec_dot))

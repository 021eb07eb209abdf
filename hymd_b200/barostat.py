"""Berendsen and stochastic-cell-rescaling barostats on top of the device-resident fields
("barostat box rescale -> hymd_ctx_set_box").

Same function names, argument order and return value ``(pm_stuff, change)`` as the reference's two
modules, selected the way ``main.py:130-134`` does it::

    from hymd_b200.barostat import berendsen, scr        # berendsen.isotropic, scr.semiisotropic, ...

=========================  ==============================
``berendsen.isotropic``    ``hymd/barostat.py:43-166``
``berendsen.semiisotropic````hymd/barostat.py:169-309``
``scr.isotropic``          ``hymd/barostat_scr.py:31-131``
``scr.semiisotropic``      ``hymd/barostat_scr.py:134-284``
=========================  ==============================

What differs from the reference: instead of building a new particle mesh with ``initialize_pm``
(``barostat.py:163-164``) the existing context is told the new box (``pm.set_box`` ->
``hymd_ctx_set_box``: new wave numbers and filter tables, same buffers and plans) and the SAME
``pm_stuff`` tuple is returned, so the field handles the caller holds stay valid.  ``positions`` may be
a torch CUDA tensor or a numpy array; it is scaled in place like the reference does.

Kept from the reference on purpose (its arithmetic is the specification): the semi-isotropic variants
scale ``positions[:][0:2]`` and ``positions[:][2]`` (``barostat.py:287, 302``), which for an ``(N, 3)``
array are the first two / the third PARTICLE rows, not the x, y / z columns; the SCR barostats return
``change = False`` from ``isotropic`` (``barostat_scr.py:131``); random numbers are drawn from the
caller's ``prng`` before the pressure is computed, in the reference's order.
"""
from __future__ import annotations

import types

import numpy as np

from .pressure import comp_pressure

BAR_PER_KJ_MOL_NM3 = 16.61


def _reinitialize(pm_stuff, config):
    """``pm_stuff = initialize_pm(pmesh, config, comm)`` without rebuilding anything."""
    pm_stuff[0].set_box(config.box_size)
    return pm_stuff


def _pressure(args, comm):
    (phi, phi_q, psi, hamiltonian, positions, velocities, config, phi_fft, phi_laplacian, phi_transfer,
     bond_pr, angle_pr) = args
    return comp_pressure(phi, phi_q, psi, hamiltonian, velocities, config, phi_fft, phi_laplacian,
                         phi_transfer, positions, bond_pr, angle_pr, comm=comm)


def _berendsen_isotropic(pmesh, pm_stuff, phi, phi_q, psi, hamiltonian, positions, velocities, config,
                         phi_fft, phi_laplacian, phi_transfer, bond_pr, angle_pr, step, prng, comm=None):
    beta = 4.6 * 10 ** (-5)                      # bar^-1, barostat.py:133
    change = False
    if np.mod(step, config.n_b) == 0:
        change = True
        pressure = _pressure((phi, phi_q, psi, hamiltonian, positions, velocities, config, phi_fft,
                              phi_laplacian, phi_transfer, bond_pr, angle_pr), comm)
        P = np.average(pressure[-3:-1]) * BAR_PER_KJ_MOL_NM3            # barostat.py:154-155
        alpha = (1.0 - (config.time_step * config.respa_inner) * config.n_b / config.tau_p * beta
                 * (config.target_pressure.P_L - P)) ** (1 / 3)
        config.box_size *= alpha
        positions *= alpha
        pm_stuff = _reinitialize(pm_stuff, config)
    return (pm_stuff, change)


def _berendsen_semiisotropic(pmesh, pm_stuff, phi, phi_q, psi, hamiltonian, positions, velocities, config,
                             phi_fft, phi_laplacian, phi_transfer, bond_pr, angle_pr, step, prng, comm=None):
    beta = 4.6 * 10 ** (-5)
    change = False
    if np.mod(step, config.n_b) == 0:
        change = True
        pressure = _pressure((phi, phi_q, psi, hamiltonian, positions, velocities, config, phi_fft,
                              phi_laplacian, phi_transfer, bond_pr, angle_pr), comm)
        PL = (pressure[-3] + pressure[-2]) / 2 * BAR_PER_KJ_MOL_NM3     # barostat.py:267-270
        PN = pressure[-1] * BAR_PER_KJ_MOL_NM3
        if config.target_pressure.P_L:
            alphaL = (1.0 - (config.time_step * config.respa_inner) * config.n_b / config.tau_p * beta
                      * (config.target_pressure.P_L - PL)) ** (1 / 3)
            config.box_size[0:2] *= alphaL
            positions[:][0:2] *= alphaL
        if config.target_pressure.P_N:
            alphaN = (1.0 - (config.time_step * config.respa_inner) * config.n_b / config.tau_p * beta
                      * (config.target_pressure.P_N - PN)) ** (1 / 3)
            config.box_size[2] *= alphaN
            positions[:][2] *= alphaN
        pm_stuff = _reinitialize(pm_stuff, config)
    return (pm_stuff, change)


def _scr_isotropic(pmesh, pm_stuff, phi, phi_q, psi, hamiltonian, positions, velocities, config,
                   phi_fft, phi_laplacian, phi_transfer, bond_pr, angle_pr, step, prng, comm=None):
    beta = 7.6 * 10 ** (-4)                      # barostat_scr.py:72
    if np.mod(step, config.n_b) == 0:
        R = prng.normal()
        pressure = _pressure((phi, phi_q, psi, hamiltonian, positions, velocities, config, phi_fft,
                              phi_laplacian, phi_transfer, bond_pr, angle_pr), comm)
        P = np.average(pressure[-3:-1]) * BAR_PER_KJ_MOL_NM3
        V = np.prod(config.box_size)
        dt = config.time_step * config.respa_inner
        noise_term = np.sqrt(2.0 * config.n_b * config.gas_constant * config.target_temperature * beta
                             * dt * config.n_b / (V * config.tau_p)) * R
        log_alpha = -config.n_b * dt * beta / config.tau_p * (config.target_pressure.P_L - P)
        alpha = np.exp((log_alpha + noise_term) / 3.0)
        config.box_size *= alpha
        positions *= alpha
        pm_stuff = _reinitialize(pm_stuff, config)
    return (pm_stuff, False)


def _scr_semiisotropic(pmesh, pm_stuff, phi, phi_q, psi, hamiltonian, positions, velocities, config,
                       phi_fft, phi_laplacian, phi_transfer, bond_pr, angle_pr, step, prng, comm=None):
    beta = 7.6 * 10 ** (-4)
    change = False
    if np.mod(step, config.n_b) == 0:
        Rxy = prng.normal()
        Rz = prng.normal()
        change = True
        pressure = _pressure((phi, phi_q, psi, hamiltonian, positions, velocities, config, phi_fft,
                              phi_laplacian, phi_transfer, bond_pr, angle_pr), comm)
        PL = (pressure[-3] + pressure[-2]) / 2 * BAR_PER_KJ_MOL_NM3
        PN = pressure[-1] * BAR_PER_KJ_MOL_NM3
        config.surface_tension = config.box_size[2] / 2 * (PN - PL)     # bar nm, barostat_scr.py:207
        dt = config.time_step * config.respa_inner
        if config.target_pressure.P_L:
            V = np.prod(config.box_size)
            noise_term = np.sqrt(4.0 * config.n_b * config.gas_constant * config.target_temperature * beta
                                 * dt * config.n_b / (3 * V * config.tau_p)) * Rxy
            log_alpha = (-2.0 * config.n_b * dt * beta / (3 * config.tau_p)
                         * (config.target_pressure.P_L - PL - config.surface_tension / config.box_size[2]))
            alpha = np.exp((log_alpha + noise_term) / 2.0)
            config.box_size[0:2] *= alpha
            positions[:][0:2] *= alpha
        if config.target_pressure.P_N:
            V = np.prod(config.box_size)
            noise_term = np.sqrt(2.0 * config.n_b * config.gas_constant * config.target_temperature * beta
                                 * dt * config.n_b / (3 * V * config.tau_p)) * Rz
            log_alpha = -config.n_b * dt * beta / (3 * config.tau_p) * (config.target_pressure.P_N - PN)
            alpha = np.exp((log_alpha + noise_term) / 1.0)
            config.box_size[2] *= alpha
            positions[:][2] *= alpha
        pm_stuff = _reinitialize(pm_stuff, config)
    return (pm_stuff, change)


berendsen = types.SimpleNamespace(isotropic=_berendsen_isotropic, semiisotropic=_berendsen_semiisotropic)
scr = types.SimpleNamespace(isotropic=_scr_isotropic, semiisotropic=_scr_semiisotropic)
# the default barostat_type is "berendsen" (input_parser.py:1052-1054)
isotropic, semiisotropic = _berendsen_isotropic, _berendsen_semiisotropic

"""CPU restatement of HyMD's CSVR thermostat and centre-of-mass momentum removal.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``hymd/thermostat.py:12-15`` (``cancel_com_momentum``) and ``hymd/thermostat.py:111-219``
(``csvr_thermostat``, Bussi-Parrinello 2008 form) on a single rank (every ``comm.allreduce`` is the
identity).  Pinned against the reference's known answers ``test/test_thermostat.py:117-143`` and
against outputs of the real ``hymd/thermostat.py`` executed in the build container
(``tests/golden/make_reference_golden.py`` -> ``tests/golden/thermostat_golden.npz``).

Two behaviours of the reference are kept on purpose:

* with ``remove_center_of_mass_momentum=False`` (or a group of one particle) the kinetic energy is
  taken over ALL particles and ALL velocities are rescaled, once per coupling group
  (``thermostat.py:190-191, 216-217`` index ``velocity[...]``, not ``velocity[ind]``);
* the random numbers are drawn in the order Gaussian, then chi-squared, per group (``:200-201``).
"""
from __future__ import annotations

import numpy as np


def cancel_com_momentum(velocities, n_particles):
    """``thermostat.py:12-15``."""
    com = np.sum(velocities, axis=0)
    return velocities - com / n_particles


def csvr_thermostat(velocity, group_of, n_groups, *, mass, gas_constant, target_temperature,
                    time_step, respa_inner, tau, draws, remove_center_of_mass_momentum=True):
    """``thermostat.py:177-219``.  ``group_of[i]`` = coupling group of particle i (-1: none);
    ``draws`` = [(R, SNf), ...] per group, the values ``random_gaussian`` / ``random_chi_squared``
    return.  Modifies ``velocity`` in place and returns the thermostat work ``sum dK``."""
    work = 0.0
    for g in range(n_groups):
        ind = np.where(group_of == g)[0]
        n_g = len(ind)
        if remove_center_of_mass_momentum and n_g > 1:
            com = np.sum(velocity[ind], axis=0)
            clean = velocity[ind] - com / n_g
            K = 0.5 * mass * np.sum(clean ** 2)
        else:
            K = 0.5 * mass * np.sum(velocity ** 2)
        K_target = 1.5 * gas_constant * n_g * target_temperature
        N_f = 3 * n_g
        c = np.exp(-(time_step * respa_inner) / tau)
        R, SNf = draws[g]
        alpha2 = (c + (1 - c) * (SNf + R ** 2) * K_target / (N_f * K)
                  + 2 * R * np.sqrt(c * (1 - c) * K_target / (N_f * K)))
        dK = K * (alpha2 - 1)
        alpha = np.sqrt(alpha2)
        if remove_center_of_mass_momentum and n_g > 1:
            clean *= alpha
            velocity[ind] = clean + com / n_g
        else:
            velocity *= alpha
        work += dK
    return work

"""numpy restatement of ``hymd/hamiltonian.py`` (filter + interaction functionals).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference builds its energy density ``w(phi)`` symbolically with sympy and
lambdifies ``w`` and ``dw/dphi_t``; here the same expressions are written out by
hand.  ``tests/golden/hamiltonian_golden.npz`` (made by ``tests/golden/
make_golden.py`` from the reference's real module) pins these against the
reference's lambdas.

* ``Hamiltonian._setup``  ``hamiltonian.py:34-71``  -> :func:`window`, rho0/a rule
* ``SquaredPhi.setup``    ``hamiltonian.py:142-204``
* ``DefaultNoChi.setup``  ``hamiltonian.py:256-319``
* ``DefaultWithChi.setup````hamiltonian.py:369-486``
* ``w_elec``              ``hamiltonian.py:148-155`` (identical in all three)
* ``V_bar_0`` / ``V_bar`` ``hamiltonian.py:157-186, 271-301, 423-475`` (pressure terms)
"""
from __future__ import annotations

import numpy as np


def window(k, sigma):
    """``exp(-sigma^2 |k|^2 / 2)`` (``hamiltonian.py:57-66``); ``k`` = 3 broadcastable arrays."""
    return np.exp(-0.5 * sigma ** 2 * (k[0] ** 2 + k[1] ** 2 + k[2] ** 2))


def setup_density_parameters(config):
    """rho0 / a bookkeeping of ``Hamiltonian._setup`` (``hamiltonian.py:39-47, 54-55``)."""
    if not hasattr(config, "simulation_volume") or config.simulation_volume is None:
        config.simulation_volume = float(np.prod(np.asarray(config.box_size)))
    if not getattr(config, "barostat", None):
        config.rho0 = config.n_particles / config.simulation_volume
        config.a = config.rho0
    if not config.rho0:
        config.rho0 = config.n_particles / config.simulation_volume
    if not getattr(config, "self_energy", None):
        config.self_energy = 0.0


def chi_matrix(config):
    """Symmetric (T,T) chi table from ``config.chi`` keyed by sorted type-name pairs
    (``hamiltonian.py:383-387, 402-407, 431-440``); zero diagonal, zero for missing pairs."""
    n = config.n_types
    chi = np.zeros((n, n), dtype=np.float64)
    table = {tuple(sorted([c.atom_1, c.atom_2])): c.interaction_energy for c in (config.chi or [])}
    for i in range(n):
        for j in range(n):
            ni, nj = config.type_to_name_map[i], config.type_to_name_map[j]
            if ni != nj:
                chi[i, j] = table.get(tuple(sorted([ni, nj])), 0.0)
    return chi


class OracleHamiltonian:
    """Holds ``H``, ``w_0``, ``v_ext[t]`` and ``w_elec`` for one of the three functionals."""

    def __init__(self, config, kind=None):
        self.config = config
        setup_density_parameters(config)
        kind = (kind or config.hamiltonian or "DefaultNoChi").lower()
        self.kind = kind
        kappa, rho0, a = config.kappa, config.rho0, config.a
        n = config.n_types
        if kind == "defaultwithchi":
            chi = chi_matrix(config)
        elif kind in ("defaultnochi", "squaredphi"):
            chi = np.zeros((n, n))
        else:
            raise ValueError(f"unknown hamiltonian {kind!r}")
        shift = 0.0 if kind == "squaredphi" else a
        self.chi = chi
        self.shift = shift

        def w_0(phi):
            total = sum(phi)
            w = 0.5 / (kappa * rho0) * (total - shift) ** 2
            for i in range(n):
                for j in range(i + 1, n):
                    if chi[i, j] != 0.0:
                        w = w + chi[i, j] * phi[i] * phi[j] / rho0
            return w

        def make_v(t):
            def v(phi):
                out = 1.0 / (kappa * rho0) * (sum(phi) - shift)
                for j in range(n):
                    if j != t and chi[t, j] != 0.0:
                        out = out + chi[t, j] * phi[j] / rho0
                return out
            return v

        def w_elec(args):
            phi_q, psi = args
            return 0.5 * phi_q * psi - config.self_energy / config.simulation_volume

        def make_vbar(t):
            # V_bar_0(phi, t) + V_bar_elec(psi, t) = V_bar_0 + type_charges[t] * psi
            v0 = make_v(t)
            return lambda args: v0(args[0]) + config.type_charges[t] * args[1]

        self.w_0 = w_0
        self.v_ext = [make_v(t) for t in range(n)]
        # V_bar_0[t] is written out separately in the reference but is the same expression as
        # dw/dphi_t for all three functionals (pinned by the golden vectors)
        self.V_bar_0 = [make_v(t) for t in range(n)]
        self.V_bar = [make_vbar(t) for t in range(n)]
        self.w_elec = w_elec

    def H(self, k, v):
        return v * window(k, self.config.sigma)

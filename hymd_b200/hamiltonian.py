"""Interaction-energy functionals of HyMD (mirror of ``hymd/hamiltonian.py``).

Same class names, constructor signatures and attributes as the reference
(``Hamiltonian``, ``SquaredPhi``, ``DefaultNoChi``, ``DefaultWithChi``,
``get_hamiltonian``; ``hamiltonian.py:16-512``): ``H``, ``w_0``, ``w``, ``w_elec``,
``v_ext[t]``, ``V_bar_0[t]``, ``V_bar[t]``.  The callables are plain closures (no sympy)
that work on numpy arrays and torch tensors alike.

The CUDA path never calls them on mesh data: every shipped functional has an AFFINE
external potential, V_t = sum_j A[t][j] phi~_j + c[t], and the fused k-space kernel takes
(A, c) directly.  :func:`affine_parameters` extracts (A, c) from any Hamiltonian object --
ours or the reference's -- by probing its ``v_ext`` callables, and rejects functionals that
are not affine.
"""
from __future__ import annotations

import math

import numpy as np


class Hamiltonian:
    """``hamiltonian.py:16-71``: filter H and the rho0 / a bookkeeping."""

    def __init__(self, config):
        self.config = config
        self._setup()

    def _setup(self):
        cfg = self.config
        if getattr(cfg, "simulation_volume", None) is None:
            # the product is taken in box_size's own dtype (float32 from the input parser,
            # input_parser.py:736), exactly like hamiltonian.py:40
            cfg.simulation_volume = float(np.prod(np.asarray(cfg.box_size)))
        if not getattr(cfg, "barostat", None):
            cfg.rho0 = cfg.n_particles / cfg.simulation_volume
            cfg.a = cfg.rho0
        if not cfg.rho0:
            cfg.rho0 = cfg.n_particles / cfg.simulation_volume
        if not getattr(cfg, "self_energy", None):
            cfg.self_energy = 0.0
        sigma = cfg.sigma

        def window(k):
            k2 = k[0] ** 2 + k[1] ** 2 + k[2] ** 2
            return np.exp(-0.5 * sigma ** 2 * k2) if not hasattr(k2, "exp") else (-0.5 * sigma ** 2 * k2).exp()

        self.window_function_lambda = window
        self.H = lambda k, v: v * window(k)

    # shared by the three functionals (hamiltonian.py:148-155, 262-269, 414-421)
    def _make_common(self, chi, shift):
        cfg = self.config
        n = cfg.n_types
        kappa, rho0 = cfg.kappa, cfg.rho0
        self.chi_matrix = chi
        self.shift = shift

        def w_0(phi):
            w = 0.5 / (kappa * rho0) * (sum(phi) - shift) ** 2
            for i in range(n):
                for j in range(i + 1, n):
                    if chi[i, j] != 0.0:
                        w = w + chi[i, j] * phi[i] * phi[j] / rho0
            return w

        def w_elec(args):
            phi_q, psi = args
            return 0.5 * phi_q * psi - cfg.self_energy / cfg.simulation_volume

        def make_v0(t):
            def v(phi):
                out = 1.0 / (kappa * rho0) * (sum(phi) - shift)
                for j in range(n):
                    if j != t and chi[t, j] != 0.0:
                        out = out + chi[t, j] * phi[j] / rho0
                return out
            return v

        def make_vbar(t):
            v0 = make_v0(t)
            return lambda args: v0(args[0]) + cfg.type_charges[t] * args[1]

        self.w_0 = w_0
        self.w_elec = w_elec
        self.v_ext = [make_v0(t) for t in range(n)]
        self.V_bar_0 = [make_v0(t) for t in range(n)]
        self.V_bar = [make_vbar(t) for t in range(n)]
        if cfg.coulombtype == "PIC_Spectral":
            self.w = lambda args: w_0(args[0]) + w_elec((args[1], args[2]))
        else:
            self.w = w_0


class SquaredPhi(Hamiltonian):
    """w = (sum phi)^2 / (2 kappa rho0)   (``hamiltonian.py:74-204``)."""

    def __init__(self, config):
        super().__init__(config)
        self._make_common(np.zeros((config.n_types, config.n_types)), 0.0)


class DefaultNoChi(Hamiltonian):
    """w = (sum phi - a)^2 / (2 kappa rho0)   (``hamiltonian.py:207-319``)."""

    def __init__(self, config):
        super().__init__(config)
        self._make_common(np.zeros((config.n_types, config.n_types)), config.a)


class DefaultWithChi(Hamiltonian):
    """w = sum_{i<j} chi_ij phi_i phi_j / rho0 + (sum phi - a)^2 / (2 kappa rho0)
    (``hamiltonian.py:322-486``)."""

    def __init__(self, config, unique_names, type_to_name_map):
        super().__init__(config)
        self.type_to_name_map = type_to_name_map
        self.chi_type_dictionary = {
            tuple(sorted([c.atom_1, c.atom_2])): c.interaction_energy for c in config.chi
        }
        n = config.n_types
        chi = np.zeros((n, n), dtype=np.float64)
        for i in range(n):
            for j in range(n):
                ni, nj = type_to_name_map[i], type_to_name_map[j]
                if ni != nj:
                    chi[i, j] = self.chi_type_dictionary[tuple(sorted([ni, nj]))]
        self._make_common(chi, config.a)


def get_hamiltonian(config):
    """``hamiltonian.py:489-512``."""
    kind = config.hamiltonian.lower()
    if kind == "defaultnochi":
        return DefaultNoChi(config)
    if kind == "defaultwithchi":
        return DefaultWithChi(config, config.unique_names, config.type_to_name_map)
    if kind == "squaredphi":
        return SquaredPhi(config)
    raise ValueError(f"unknown hamiltonian {config.hamiltonian!r}")


def affine_parameters(hamiltonian, n_types):
    """(A, c) with v_ext[t](phi) == sum_j A[t,j] phi_j + c[t], obtained by probing the
    Hamiltonian's own ``v_ext`` callables (works for the reference's sympy lambdas too).
    Raises ValueError for a non-affine functional."""
    T = n_types
    zero = [np.float64(0.0)] * T
    c = np.array([float(hamiltonian.v_ext[t](zero)) for t in range(T)])
    A = np.zeros((T, T))
    for j in range(T):
        e = [np.float64(1.0 if i == j else 0.0) for i in range(T)]
        for t in range(T):
            A[t, j] = float(hamiltonian.v_ext[t](e)) - c[t]
    rng = np.random.default_rng(12345)
    for _ in range(2):
        x = rng.uniform(0.1, 3.0, size=T)
        for t in range(T):
            got = float(hamiltonian.v_ext[t](list(x)))
            want = float(A[t] @ x + c[t])
            if not math.isclose(got, want, rel_tol=1e-9, abs_tol=1e-9 * (abs(c[t]) + np.abs(A[t]).sum())):
                raise ValueError(
                    "hymd_b200 supports Hamiltonians with an affine external potential "
                    f"(SquaredPhi, DefaultNoChi, DefaultWithChi); v_ext[{t}] is not affine")
    return A, c


def energy_parameters(hamiltonian):
    """(chi[T,T], kappa, rho0, a_shift) for the device energy reduction, or None when the
    functional is not one of the three known kinds (the caller then evaluates
    ``hamiltonian.w_0`` on device tensors)."""
    cfg = hamiltonian.config
    name = type(hamiltonian).__name__
    T = cfg.n_types
    if name == "SquaredPhi":
        return np.zeros((T, T)), float(cfg.kappa), float(cfg.rho0), 0.0
    if name == "DefaultNoChi":
        return np.zeros((T, T)), float(cfg.kappa), float(cfg.rho0), float(cfg.a)
    if name == "DefaultWithChi":
        chi = np.zeros((T, T))
        table = hamiltonian.chi_type_dictionary
        names = hamiltonian.type_to_name_map
        for i in range(T):
            for j in range(T):
                if names[i] != names[j]:
                    chi[i, j] = float(table[tuple(sorted([names[i], names[j]]))])
        return chi, float(cfg.kappa), float(cfg.rho0), float(cfg.a)
    return None

"""Closed-form / direct-sum known answers for the field-force cycle.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

* Gaussian-core energy: the closed form the reference's own tests assert against
  (``test/test_hamiltonian.py:213-227, 334-360``).
* Gaussian-core pair forces: ``hymd/gaussian_core.py:34-54`` /
  ``hymd/compute_gaussian_core.f90:45-85`` (the reference's dev-time force oracle).
* Gaussian-smeared reciprocal-space Ewald sum over the mesh's own wave vectors
  (probe 3) for the PME branch ``field.py:356-403``.
"""
from __future__ import annotations

import numpy as np


def gaussian_core_energy(r, type_ids, chi, kappa, sigma, rho0, with_a_shift=True):
    """E = N/c + sum_{i<j} 2 (1 + kappa chi_ij) exp(-r_ij^2 / 4 sigma^2)/c [- N/(2 kappa)]."""
    r = np.asarray(r, dtype=np.float64)
    n = r.shape[0]
    c = 16.0 * np.pi ** 1.5 * kappa * sigma ** 3 * rho0
    e = n / c
    for i in range(n):
        for j in range(i + 1, n):
            d = r[i] - r[j]
            g = np.exp(-np.dot(d, d) / (4.0 * sigma ** 2))
            e += 2.0 * (1.0 + kappa * chi[type_ids[i], type_ids[j]]) * g / c
    if with_a_shift:
        e -= 0.5 * n / kappa
    return e


def gaussian_core_forces(r, type_ids, chi, kappa, sigma, rho0, box=None, images=0):
    """F_i = sum_j r_ij exp(-r_ij^2/4 sigma^2)(1 + kappa chi_ij)/(16 pi^1.5 kappa sigma^5 rho0),
    r_ij = r_i - r_j (repulsive for chi >= 0); optional periodic images."""
    r = np.asarray(r, dtype=np.float64)
    n = r.shape[0]
    pref = 1.0 / (16.0 * np.pi ** 1.5 * kappa * sigma ** 5 * rho0)
    f = np.zeros_like(r)
    shifts = [np.zeros(3)]
    if box is not None and images > 0:
        rng = range(-images, images + 1)
        shifts = [np.array([a, b, c]) * np.asarray(box, dtype=np.float64)
                  for a in rng for b in rng for c in rng]
    for i in range(n):
        for j in range(n):
            for s in shifts:
                if i == j and not np.any(s):
                    continue
                d = r[i] - r[j] - s
                g = np.exp(-np.dot(d, d) / (4.0 * sigma ** 2))
                f[i] += pref * (1.0 + kappa * chi[type_ids[i], type_ids[j]]) * g * d
    return f


def ewald_reciprocal_on_mesh_k(q, r, mesh, box, sigma, conv):
    """Direct structure-factor sum over the mesh's wave vectors.

    psi(r) = (4 pi conv / V) sum_{k != 0} H(k) S(k) e^{ik.r} / k^2 ,  S(k) = sum_j q_j e^{-ik.r_j}
    E = 1/2 sum_i q_i psi(r_i) ;  F_i = -q_i grad psi(r_i)
    Odd mesh sizes only (no Nyquist ambiguity).  Returns (energy_without_self_term, forces).
    """
    q = np.asarray(q, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    box = np.asarray(box, dtype=np.float64)
    assert all(n % 2 == 1 for n in mesh)
    ks = [2.0 * np.pi * np.fft.fftfreq(n, d=l / n) for n, l in zip(mesh, box)]
    kx, ky, kz = np.meshgrid(*ks, indexing="ij")
    k2 = kx ** 2 + ky ** 2 + kz ** 2
    kvec = np.stack([kx.ravel(), ky.ravel(), kz.ravel()], axis=1)
    k2 = k2.ravel()
    nz = k2 > 0
    kvec, k2 = kvec[nz], k2[nz]
    h = np.exp(-0.5 * sigma ** 2 * k2)
    phase = np.exp(-1j * (r @ kvec.T))              # (N, K): e^{-ik.r_j}
    s = q @ phase                                   # S(k)
    vol = float(np.prod(box))
    coef = 4.0 * np.pi * conv / vol * h / k2
    energy = 0.5 * float(np.sum(coef * np.abs(s) ** 2))
    # psi(r_i) = sum_k coef S(k) e^{+ik.r_i};  grad -> i k
    grad = np.real((np.conj(phase) * (coef * s)[None, :] * 1j) @ kvec)   # (N,3)
    forces = -q[:, None] * grad
    return energy, forces

"""CPU restatement of HyMD's intramolecular (bonded) force kernels.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates, term by term and in float64 like the Fortran (``real(8)`` locals), the production kernels

* ``hymd/compute_bond_forces.f90:1-61``      ``cbf``  harmonic two-particle bonds
* ``hymd/compute_angle_forces.f90:1-93``     ``caf``  harmonic three-particle angles
* ``hymd/compute_dihedral_forces.f90:1-137`` ``cdf``  dihedrals, ``dtype`` 0 (cosine series + coil
  series, ``dipole_reconstruction.f90:37-48``) and ``dtype`` 2 (improper, harmonic in phi).
  ``dtype`` 1 (combined bending-torsion with dipole reconstruction, ``dipole_reconstruction.f90:50-221``)
  belongs to the protein-dipole electrostatics and raises.

Pinned (``tests/test_oracle_bonded.py``) against the known answers of the reference's
``test/test_force.py:52-103, 135-198, 231-273`` and against outputs of the reference's own
``compute_*_forces__plain`` functions (``hymd/force.py:731-852``) executed in the build container
(``tests/golden/make_reference_golden.py`` -> ``tests/golden/bonded_golden.npz``).

Term order: forces are accumulated in term order like the Fortran loops; the pressure by-products
``bond_pr`` / ``angle_pr`` follow ``compute_bond_forces.f90:58`` / ``compute_angle_forces.f90:90``
(``bond_pr`` starts from zero here; the Fortran leaves its ``intent(out)`` accumulator
uninitialised).
"""
from __future__ import annotations

import numpy as np


def _mic(d, box):
    """``d - box * nint(d / box)`` (Fortran ``nint`` rounds half away from zero).  ``d`` arrives as the
    difference of two positions IN THE POSITION DTYPE (``r`` is ``real(4)`` in the default build, so
    ``r(bb,:) - r(aa,:)`` is a single-precision subtraction) and is widened to float64 here, like the
    assignment to the ``real(8)`` local does."""
    d = np.asarray(d, dtype=np.float64)
    q = d / box
    return d - box * (np.sign(q) * np.floor(np.abs(q) + 0.5))


def compute_bond_forces(r, box, a, b, r0, k):
    """``cbf``: returns ``(f (N,3) float64, energy, bond_pr (3,))``."""
    r = np.asarray(r)
    box = np.asarray(box, dtype=np.float64)
    f = np.zeros(r.shape, dtype=np.float64)
    energy = 0.0
    pr = np.zeros(3)
    for aa, bb, r0i, ki in zip(a, b, r0, k):
        rab = _mic(r[bb] - r[aa], box)
        n = np.sqrt(np.dot(rab, rab))
        df = ki * (n - r0i)
        fa = -df * rab / n
        f[aa] -= fa
        f[bb] += fa
        energy += 0.5 * ki * (n - r0i) ** 2
        pr += fa * rab
    return f, energy, pr


def compute_angle_forces(r, box, a, b, c, t0, k):
    """``caf``: returns ``(f, energy, angle_pr)``."""
    r = np.asarray(r)
    box = np.asarray(box, dtype=np.float64)
    f = np.zeros(r.shape, dtype=np.float64)
    energy = 0.0
    pr = np.zeros(3)
    for aa, bb, cc, t0i, ki in zip(a, b, c, t0, k):
        ra = _mic(r[aa] - r[bb], box)
        rc = _mic(r[cc] - r[bb], box)
        na = np.sqrt(np.dot(ra, ra))
        nc = np.sqrt(np.dot(rc, rc))
        ea, ec = ra / na, rc / nc
        cosphi = np.dot(ea, ec)
        if cosphi * cosphi < 1.0:
            theta = np.arccos(cosphi)
            sinphi = np.sin(theta)
            d = theta - t0i
            ff = ki * d
            xra = -ff / (na * sinphi)
            xrc = -ff / (nc * sinphi)
            fa = (ec - cosphi * ea) * xra
            fc = (ea - cosphi * ec) * xrc
            f[aa] -= fa
            f[cc] -= fc
            f[bb] += fa + fc
            energy += 0.5 * ff * d
            pr += -(fa * ra) - (fc * rc)
    return f, energy, pr


def _cosine_series(c_n, d_n, phi):
    """``dipole_reconstruction.f90:37-48``: returns (energy, dE/dphi) increments."""
    i = np.arange(len(c_n))
    e = float(np.sum(c_n * (1.0 + np.cos(i * phi - d_n))))
    de = float(-np.sum(i * c_n * np.sin(i * phi - d_n)))
    return e, de


def compute_dihedral_forces(r, box, a, b, c, d, coeff, dtype):
    """``cdf`` for ``dtype`` 0 and 2: returns ``(f, energy)``.  ``coeff`` is the (D,6,5) array that
    ``prepare_bonds`` builds (``force.py:678-690``)."""
    r = np.asarray(r)
    box = np.asarray(box, dtype=np.float64)
    coeff = np.asarray(coeff, dtype=np.float64)
    force = np.zeros(r.shape, dtype=np.float64)
    energy = 0.0
    for ind, (aa, bb, cc, dd) in enumerate(zip(a, b, c, d)):
        f = _mic(r[aa] - r[bb], box)
        g = _mic(r[bb] - r[cc], box)
        h = _mic(r[dd] - r[cc], box)
        v = np.cross(f, g)
        w = np.cross(h, g)
        v_sq = np.dot(v, v)
        w_sq = np.dot(w, w)
        g_norm = np.sqrt(np.dot(g, g))
        cos_phi = np.dot(v, w)
        sin_phi = np.dot(w, f) * g_norm
        phi = np.arctan2(sin_phi, cos_phi)
        f_dot_g = np.dot(f, g)
        h_dot_g = np.dot(h, g)
        df = 0.0
        if dtype[ind] == 0:
            e, de = _cosine_series(coeff[ind, 0], coeff[ind, 1], phi)
            energy += e
            df += de
            c_coil, d_coil = coeff[ind, 2], coeff[ind, 3]
            if np.any(c_coil != 0) and np.any(d_coil != 0):
                e, de = _cosine_series(c_coil, d_coil, phi)
                energy += e
                df += de
        elif dtype[ind] == 2:
            eq_value, force_const = coeff[ind, 0, 0], coeff[ind, 0, 1]
            df = force_const * (phi - eq_value)
            energy += 0.5 * force_const * (phi - eq_value) ** 2
        else:
            raise NotImplementedError("dihedral dtype 1 (CBT + dipole reconstruction) is row f3")
        sc = v * f_dot_g / (v_sq * g_norm) - w * h_dot_g / (w_sq * g_norm)
        fa = -df * g_norm * v / v_sq
        fd = df * g_norm * w / w_sq
        fb = df * sc - fa
        fc = -df * sc - fd
        force[aa] += fa
        force[bb] += fb
        force[cc] += fc
        force[dd] += fd
    return force, energy


def dihedral_energy(r, box, a, b, c, d, coeff, dtype):
    return compute_dihedral_forces(r, box, a, b, c, d, coeff, dtype)[1]

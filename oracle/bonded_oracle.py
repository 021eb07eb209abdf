"""CPU restatement of HyMD's intramolecular (bonded) force kernels.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates, term by term and in float64 like the Fortran (``real(8)`` locals), the production kernels

* ``hymd/compute_bond_forces.f90:1-61``      ``cbf``  harmonic two-particle bonds
* ``hymd/compute_angle_forces.f90:1-93``     ``caf``  harmonic three-particle angles
* ``hymd/compute_dihedral_forces.f90:1-137`` ``cdf``  dihedrals, ``dtype`` 0 (cosine series + coil
  series, ``dipole_reconstruction.f90:37-48``), ``dtype`` 2 (improper, harmonic in phi) and ``dtype`` 1
  (combined bending-torsion: the same series plus the ``reconstruct`` routine,
  ``dipole_reconstruction.f90:50-221``, which also places the backbone dipoles and builds the transfer
  matrices that ``dipole_forces_redistribution``, ``hymd/force.py:855-880``, applies).
  **Parity unpinned for dtype 1**: no reference test exercises it (``grep -rn dipole test/`` finds the CLI
  switch only) and there is no Fortran compiler in the build container, so this part of the restatement is
  checked for internal consistency only (``tests/test_oracle_bonded.py``: forces = -grad E and transfer
  matrices = d(dipole)/d(bead) by finite differences).

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


# ``real(8), parameter :: ... cos_psi = cos(1.392947), sin_psi = sin(1.392947)``: default-real literals, so the
# Fortran evaluates both in single precision and widens (dipole_reconstruction.f90:76)
_DELTA = 0.3
_COS_PSI = float(np.cos(np.float32(1.392947), dtype=np.float32))
_SIN_PSI = float(np.sin(np.float32(1.392947), dtype=np.float32))


def _cross_matrix(m, v):
    """Row i of the result = row i of ``m`` x ``v`` (dipole_reconstruction.f90:14-24)."""
    return np.cross(m, v[None, :])


def reconstruct(rab, rb, rcb, box, c_k, d_k, phi, dipole_flag, gamma_sign=1.0):
    """``reconstruct`` (dipole_reconstruction.f90:50-221) for the angle a-b-c with ``rab = r_a - r_b`` and
    ``rcb = r_c - r_b``: returns ``(energy_cbt, df_cbt, fa, fb, fc, dipole (2,3) | None, transfer (3,3,3) | None)``
    or ``None`` for collinear bonds (``cos^2 gamma >= 1``: the Fortran leaves its outputs untouched).

    ``gamma_sign``: the Fortran's ``fa, fb, fc`` ARE ``+d gamma / d r_i`` (checked by finite differences), but
    its transfer matrices use them as if they were ``-d gamma / d r_i`` in the three places where the chain rule
    goes through gamma (``cos_gamma * outer_product(f, n)`` in ``N_i``, ``FN_i``, ``FM_i``, lines 186-206).  The
    default +1 is the reference as written (what parity means); with -1 the matrices are the exact Jacobians
    ``d d / d r_i`` of the half dipole vector at fixed phi, which ``tests/test_oracle_bonded.py`` uses to check
    every other term of this restatement against finite differences."""
    k, dk = _cosine_series(c_k, d_k, phi)
    gamma_0 = 1.85 - 0.227 * np.cos(phi - 0.785)
    dg = 0.227 * np.sin(phi - 0.785)
    norm_a = np.sqrt(np.dot(rab, rab))
    norm_c = np.sqrt(np.dot(rcb, rcb))
    w = rab / norm_a
    v = rcb / norm_c
    cos_gamma = float(np.dot(w, v))
    cos2 = cos_gamma * cos_gamma
    if not cos2 < 1.0:
        return None
    gamm = np.arccos(cos_gamma)
    sin_gamma = np.sqrt(1.0 - cos2)
    if sin_gamma < 0.1:
        sin_gamma = float(np.float32(0.1))      # "sin_gamma = 0.1": a default-real literal
    fa = -((v - cos_gamma * w) / norm_a) / sin_gamma
    fc = -((w - cos_gamma * v) / norm_c) / sin_gamma
    fb = -(fa + fc)
    df_ang = k * (gamm - gamma_0)
    var_sq = (gamm - gamma_0) ** 2
    energy_cbt = 0.5 * k * var_sq
    df_cbt = 0.5 * dk * var_sq - df_ang * dg
    if dipole_flag == 0:
        return energy_cbt, df_cbt, df_ang * fa, df_ang * fb, df_ang * fc, None, None
    fac = np.exp((gamm - 1.73) / 0.025)
    theta = -1.607 * gamm + 0.094 + 1.883 / (1.0 + fac)
    d_theta = -1.607 - 1.883 / 0.025 * fac / ((1.0 + fac) ** 2)
    cos_theta, sin_theta = np.cos(theta), np.sin(theta)
    n = np.cross(w, v) / sin_gamma
    m = np.cross(n, v)
    r0 = np.asarray(rb, dtype=np.float64) + 0.5 * rcb
    dvec = 0.5 * _DELTA * (_COS_PSI * v + _SIN_PSI * (cos_theta * n + sin_theta * m))
    dipole = np.stack([r0 + dvec, r0 - dvec])
    # dipole(i,:) is real(4): the position is rounded to the array's precision before the wrap
    eye = np.eye(3)
    V_b = (np.outer(v, v) - eye) / norm_c
    W_b = (np.outer(w, w) - eye) / norm_a
    V_c, W_a = -V_b, -W_b
    gs = float(gamma_sign)
    N_a = (gs * cos_gamma * np.outer(fa, n) + _cross_matrix(W_a, v)) / sin_gamma
    N_b = (gs * cos_gamma * np.outer(fb, n) + _cross_matrix(W_b, v) - _cross_matrix(V_b, w)) / sin_gamma
    N_c = (gs * cos_gamma * np.outer(fc, n) - _cross_matrix(V_c, w)) / sin_gamma
    M_a = _cross_matrix(N_a, v)
    M_b = _cross_matrix(N_b, v) - _cross_matrix(V_b, n)
    M_c = _cross_matrix(N_c, v) - _cross_matrix(V_c, n)
    FN = [gs * sin_theta * d_theta * np.outer(x, n) for x in (fa, fb, fc)]
    FM = [gs * cos_theta * d_theta * np.outer(x, m) for x in (fa, fb, fc)]
    T = np.zeros((3, 3, 3))
    T[0] = 0.5 * _DELTA * (_SIN_PSI * (cos_theta * N_a + sin_theta * M_a + FN[0] - FM[0]))
    T[1] = 0.5 * _DELTA * (_COS_PSI * V_b + _SIN_PSI * (cos_theta * N_b + sin_theta * M_b + FN[1] - FM[1]))
    T[2] = 0.5 * _DELTA * (_COS_PSI * V_c + _SIN_PSI * (cos_theta * N_c + sin_theta * M_c + FN[2] - FM[2]))
    return energy_cbt, df_cbt, df_ang * fa, df_ang * fb, df_ang * fc, dipole, T


def compute_dihedral_forces(r, box, a, b, c, d, coeff, dtype, bb_index=None, dipole_flag=0, full=False):
    """``cdf``: returns ``(f, energy)``, or with ``full`` ``(f, energy, dipoles (D,4,3), transfer (D,6,3,3))`` in the
    dtype of ``r`` like the Fortran's ``real(4)`` arrays.  ``coeff`` is the (D,6,5) array that ``prepare_bonds``
    builds (``force.py:678-690``); ``bb_index[i] == 1`` marks the last dihedral of a backbone, whose second angle
    b-c-d is reconstructed as well (``compute_dihedral_forces.f90:97-111``)."""
    r = np.asarray(r)
    box = np.asarray(box, dtype=np.float64)
    coeff = np.asarray(coeff, dtype=np.float64)
    force = np.zeros(r.shape, dtype=np.float64)
    n4 = len(a)
    dipoles = np.zeros((n4, 4, 3), dtype=r.dtype)
    transfer = np.zeros((n4, 6, 3, 3), dtype=r.dtype)
    if bb_index is None:
        bb_index = np.zeros(n4, dtype=int)
    energy = 0.0

    def wrap(x):          # dipole(i,:) = dipole(i,:) - box * nint(dipole(i,:) / box), on the real(4) element
        x = x.astype(r.dtype).astype(np.float64)
        q = x / box
        return (x - box * (np.sign(q) * np.floor(np.abs(q) + 0.5))).astype(r.dtype)

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
        if dtype[ind] in (0, 1):
            e, de = _cosine_series(coeff[ind, 0], coeff[ind, 1], phi)
            energy += e
            df += de
            c_coil, d_coil = coeff[ind, 2], coeff[ind, 3]
            if np.any(c_coil != 0) and np.any(d_coil != 0):
                e, de = _cosine_series(c_coil, d_coil, phi)
                energy += e
                df += de
        if dtype[ind] == 1:
            c_k, d_k = coeff[ind, 4], coeff[ind, 5]
            rec = reconstruct(f, r[bb], -g, box, c_k, d_k, phi, dipole_flag)
            if rec is not None:
                e_cbt, df_cbt, fa, fb, fc, dip, tm = rec
                energy += e_cbt
                df += df_cbt
                force[aa] -= fa
                force[bb] -= fb
                force[cc] -= fc
                if dip is not None:
                    dipoles[ind, 0:2] = wrap(dip)
                    transfer[ind, 0:3] = tm
            if bb_index[ind] == 1:
                rec = reconstruct(g, r[cc], h, box, c_k, d_k, phi, dipole_flag)
                if rec is not None:
                    e_cbt, df_cbt, fb, fc, fd, dip, tm = rec
                    energy += e_cbt
                    df += df_cbt
                    force[bb] -= fb
                    force[cc] -= fc
                    force[dd] -= fd
                    if dip is not None:
                        dipoles[ind, 2:4] = wrap(dip)
                        transfer[ind, 3:6] = tm
        elif dtype[ind] == 2:
            eq_value, force_const = coeff[ind, 0, 0], coeff[ind, 0, 1]
            df = force_const * (phi - eq_value)
            energy += 0.5 * force_const * (phi - eq_value) ** 2
        elif dtype[ind] != 0:
            raise ValueError(f"dihedral dtype {dtype[ind]}")
        sc = v * f_dot_g / (v_sq * g_norm) - w * h_dot_g / (w_sq * g_norm)
        fa = -df * g_norm * v / v_sq
        fd = df * g_norm * w / w_sq
        fb = df * sc - fa
        fc = -df * sc - fd
        force[aa] += fa
        force[bb] += fb
        force[cc] += fc
        force[dd] += fd
    if full:
        return force, energy, dipoles, transfer
    return force, energy


def dipole_forces_redistribution(n_particles, f_dipoles, transfer, a, b, c, d, dtype, last):
    """``dipole_forces_redistribution`` (``hymd/force.py:855-880``): forces on the reconstructed dipole charges
    ``f_dipoles (D,4,3)`` carried back to the backbone beads through the transfer matrices; returns ``(N,3)``."""
    f_dipoles = np.asarray(f_dipoles, dtype=np.float64)
    transfer = np.asarray(transfer, dtype=np.float64)
    out = np.zeros((n_particles, 3))
    for i, j, k, l, fd, m, ty, is_last in zip(a, b, c, d, f_dipoles, transfer, dtype, last):
        if ty != 1:
            continue
        s, df = fd[0] + fd[1], fd[0] - fd[1]
        out[i] += m[0] @ df
        out[j] += m[1] @ df + 0.5 * s
        out[k] += m[2] @ df + 0.5 * s
        if is_last == 1:
            s, df = fd[2] + fd[3], fd[2] - fd[3]
            out[j] += m[3] @ df
            out[k] += m[4] @ df + 0.5 * s
            out[l] += m[5] @ df + 0.5 * s
    return out


def dihedral_energy(r, box, a, b, c, d, coeff, dtype, bb_index=None):
    return compute_dihedral_forces(r, box, a, b, c, d, coeff, dtype, bb_index)[1]

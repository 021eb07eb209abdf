"""Bonded-force oracle (oracle/bonded_oracle.py) pinned on the reference:
* known answers of the reference's test/test_force.py (a sample typed in below, the rest through
  tests/golden/bonded_golden.npz, which holds the per-term outputs of the reference's own
  ``compute_*_forces__plain`` run by tests/golden/make_reference_golden.py),
* whole-list outputs of those functions on seeded periodic chain systems,
* F = -grad E by central differences, and translation invariance across the periodic box."""
import os

import numpy as np
import pytest

from oracle import bonded_oracle as bo

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bonded_golden.npz"))


def test_bond_kats_of_reference_test_force():
    """test/test_force.py:52-103 (first and fourth bond of the DPPC molecule)."""
    r, box = G["dppc/r"], G["dppc/box"]
    a, b, r0, k = (G["dppc/b2_" + x] for x in ("a", "b", "r0", "k"))
    expected = {
        (0, 1): (0.24545803261508981, [20.998021457611852, 9.0071937483622282, -9.5707176942411820]),
        (2, 4): (9.3338621118512890, [-52.053439897669733, 83.19032187729123, 117.06607117521537]),
    }
    seen = 0
    for i in range(len(a)):
        key = (int(a[i]), int(b[i]))
        f, e, _ = bo.compute_bond_forces(r, box, a[i:i + 1], b[i:i + 1], r0[i:i + 1], k[i:i + 1])
        assert e == pytest.approx(G["dppc/b2_term_energy"][i], abs=1e-13)
        np.testing.assert_allclose(f, G["dppc/b2_term_force"][i], rtol=0, atol=1e-12)
        if key in expected:
            seen += 1
            assert e == pytest.approx(expected[key][0], abs=1e-13)
            np.testing.assert_allclose(f[key[0]], expected[key][1], rtol=0, atol=1e-12)
            np.testing.assert_allclose(f[key[1]], -np.array(expected[key][1]), rtol=0, atol=1e-12)
    assert seen == 2 and len(a) == 11                       # test_force.py:29


def test_angle_kats_of_reference_test_force():
    """test/test_force.py:135-198 (first angle P-G-G)."""
    r, box = G["dppc/r"], G["dppc/box"]
    a, b, c, t0, k = (G["dppc/b3_" + x] for x in ("a", "b", "c", "t0", "k"))
    assert len(a) == 8                                      # test_force.py:116
    for i in range(len(a)):
        f, e, _ = bo.compute_angle_forces(r, box, a[i:i + 1], b[i:i + 1], c[i:i + 1], t0[i:i + 1], k[i:i + 1])
        assert e == pytest.approx(G["dppc/b3_term_energy"][i], abs=1e-12)
        np.testing.assert_allclose(f, G["dppc/b3_term_force"][i], rtol=0, atol=1e-11)
        if (a[i], b[i], c[i]) == (1, 2, 3):
            assert e == pytest.approx(0.24138227262192161, abs=1e-12)
            np.testing.assert_allclose(f[1], [2.4096577139753332, 4.6682763444497457, 6.0136584922358995], atol=1e-11)
            np.testing.assert_allclose(f[2], [-11.393593064392608, -11.244884057123880, -6.4084691231053164], atol=1e-11)
            np.testing.assert_allclose(f[3], [8.9839353504172745, 6.5766077126741331, 0.39481063086941653], atol=1e-11)


def test_dihedral_kats_of_reference_test_force():
    """test/test_force.py:231-273.  The production Fortran kernel (compute_dihedral_forces.f90:121-134)
    returns the negative energy gradient; the deprecated ``compute_dihedral_forces__plain`` that the
    reference test exercises returns the opposite sign (force.py:845-851 vs the Fortran), so the
    oracle must equal MINUS the reference's expected forces, and the energies must agree."""
    r, box = G["ala/r"], G["ala/box"]
    a, b, c, d, coeff, dt = (G["ala/" + x] for x in ("a", "b", "c", "d", "coeff", "dtype"))
    assert len(a) == 5                                      # test_force.py:209
    exp_e0 = 5.512306711980792
    exp_fi0 = np.array([4.167404131528236, 5.964780465237133, 6.027290456833508])
    for i in range(len(a)):
        f, e = bo.compute_dihedral_forces(r, box, a[i:i + 1], b[i:i + 1], c[i:i + 1], d[i:i + 1],
                                          coeff[i:i + 1], dt[i:i + 1])
        assert e == pytest.approx(G["ala/term_energy"][i], abs=1e-12)
        np.testing.assert_allclose(f, -G["ala/term_force_plain"][i], rtol=0, atol=1e-11)
        if i == 0:
            assert e == pytest.approx(exp_e0, abs=1e-12)
            np.testing.assert_allclose(f[a[0]], -exp_fi0, rtol=0, atol=1e-11)


def test_whole_lists_match_reference_plain_functions():
    r, box = G["chains/r"], G["chains/box"]
    f, e, _ = bo.compute_bond_forces(r, box, G["chains/b2_a"], G["chains/b2_b"], G["chains/b2_r0"], G["chains/b2_k"])
    assert e == pytest.approx(float(G["chains/b2_energy"]), rel=1e-13)
    np.testing.assert_allclose(f, G["chains/b2_force"], rtol=0, atol=1e-10 * np.abs(f).max())
    f, e, _ = bo.compute_angle_forces(r, box, G["chains/b3_a"], G["chains/b3_b"], G["chains/b3_c"],
                                      G["chains/b3_t0"], G["chains/b3_k"])
    assert e == pytest.approx(float(G["chains/b3_energy"]), rel=1e-12)
    np.testing.assert_allclose(f, G["chains/b3_force"], rtol=0, atol=1e-9 * np.abs(f).max())
    r, box = G["dih/r"], G["dih/box"]
    n4 = len(G["dih/a"])
    f, e = bo.compute_dihedral_forces(r, box, G["dih/a"], G["dih/b"], G["dih/c"], G["dih/d"],
                                      G["dih/coeff"], np.zeros(n4, dtype=int))
    assert e == pytest.approx(float(G["dih/energy"]), rel=1e-12)
    np.testing.assert_allclose(f, -G["dih/force_plain"], rtol=0, atol=1e-10 * np.abs(f).max())


def _num_grad(energy, r, h=1e-6):
    g = np.zeros_like(r)
    for i in range(r.shape[0]):
        for k in range(3):
            rp, rm = r.copy(), r.copy()
            rp[i, k] += h
            rm[i, k] -= h
            g[i, k] = (energy(rp) - energy(rm)) / (2 * h)
    return g


def test_forces_are_negative_energy_gradients_and_periodic():
    rng = np.random.default_rng(7)
    box = np.array([2.0, 2.5, 3.0])
    r = np.cumsum(rng.normal(scale=0.3, size=(8, 3)), axis=0) + 1.0
    a4 = np.arange(5)
    coeff = np.zeros((5, 6, 5))
    coeff[:, 0] = rng.normal(size=(5, 5))
    coeff[:, 1] = rng.normal(size=(5, 5))
    coeff[:, 2] = rng.normal(size=(5, 5))          # coil series active (both rows non-zero)
    coeff[:, 3] = rng.normal(size=(5, 5))
    dt = np.array([0, 0, 2, 0, 2])
    coeff[2, 0, :2] = [0.4, 30.0]
    coeff[4, 0, :2] = [-1.0, 12.0]
    a3 = np.arange(6)
    args2 = (np.arange(7), np.arange(7) + 1, np.full(7, 0.4), np.full(7, 900.0))
    args3 = (a3, a3 + 1, a3 + 2, np.full(6, 2.0), np.full(6, 30.0))
    args4 = (a4, a4 + 1, a4 + 2, a4 + 3, coeff, dt)
    for fn, args, e_idx in ((bo.compute_bond_forces, args2, 1), (bo.compute_angle_forces, args3, 1),
                            (bo.compute_dihedral_forces, args4, 1)):
        f = fn(r, box, *args)[0]
        g = _num_grad(lambda x: fn(x, box, *args)[e_idx], r)
        np.testing.assert_allclose(f, -g, rtol=0, atol=2e-6 * max(1.0, np.abs(f).max()))
        # wrapping every particle into the box (different images per particle) changes nothing
        f_wrapped = fn(np.mod(r + np.array([0.9, 1.7, 2.6]), box), box, *args)[0]
        np.testing.assert_allclose(f_wrapped, f, rtol=0, atol=1e-9 * max(1.0, np.abs(f).max()))
        assert np.abs(f.sum(axis=0)).max() < 1e-9 * max(1.0, np.abs(f).max())


def test_pressure_byproducts_are_virials():
    rng = np.random.default_rng(8)
    box = np.array([3.0, 3.0, 3.0])
    r = rng.random((6, 3)) + 1.0
    a = np.arange(5)
    f, e, pr = bo.compute_bond_forces(r, box, a, a + 1, np.full(5, 0.3), np.full(5, 100.0))
    # bond_pr = sum_terms fa * rab = sum_i r_i * f_i (component-wise virial) for an open chain
    np.testing.assert_allclose(pr, np.sum(r * f, axis=0), rtol=1e-12, atol=1e-12)
    a3 = np.arange(4)
    f, e, pr = bo.compute_angle_forces(r, box, a3, a3 + 1, a3 + 2, np.full(4, 2.0), np.full(4, 25.0))
    np.testing.assert_allclose(pr, np.sum(r * f, axis=0), rtol=1e-10, atol=1e-10)


def _cbt_system(seed=11, n=7):
    """A short backbone: n beads, n-3 dihedrals of dtype 1, the last one flagged (its second angle is reconstructed)."""
    rng = np.random.default_rng(seed)
    r = np.cumsum(rng.normal(scale=0.33, size=(n, 3)), axis=0) + np.array([2.0, 2.5, 3.0])
    box = np.array([5.0, 6.0, 7.0])
    a = np.arange(n - 3)
    coeff = np.zeros((n - 3, 6, 5))
    coeff[:, 0] = rng.uniform(0.5, 2.0, size=(n - 3, 5))
    coeff[:, 1] = rng.uniform(-1, 1, size=(n - 3, 5))
    coeff[:, 4] = rng.uniform(20, 60, size=(n - 3, 5))       # k(phi) series of the bending constant
    coeff[:, 5] = rng.uniform(-1, 1, size=(n - 3, 5))
    dt = np.ones(n - 3, dtype=int)
    last = np.zeros(n - 3, dtype=int)
    last[-1] = 1
    return r, box, (a, a + 1, a + 2, a + 3), coeff, dt, last


def test_cbt_forces_are_negative_energy_gradients():
    """dtype 1 (compute_dihedral_forces.f90:77-112 + reconstruct): propensity series + k(phi) (gamma - gamma_0(phi))^2
    on the angle a-b-c of every dihedral and on b-c-d of the last one; the forces (dihedral part with the extra
    dE/dphi of the bending term, minus the angle parts) are the negative gradient of the returned energy."""
    r, box, idx, coeff, dt, last = _cbt_system()
    f, e = bo.compute_dihedral_forces(r, box, *idx, coeff, dt, last)
    g = _num_grad(lambda x: bo.compute_dihedral_forces(x, box, *idx, coeff, dt, last)[1], r)
    np.testing.assert_allclose(f, -g, rtol=0, atol=2e-6 * np.abs(f).max())
    assert np.abs(f.sum(axis=0)).max() < 1e-9 * np.abs(f).max()
    # the flag of the last dihedral matters, and dipole_flag does not change forces or energy
    f0, e0 = bo.compute_dihedral_forces(r, box, *idx, coeff, dt, np.zeros_like(last))
    assert abs(e - e0) > 1e-3
    f1, e1, dip, tm = bo.compute_dihedral_forces(r, box, *idx, coeff, dt, last, dipole_flag=1, full=True)
    assert e1 == e and np.array_equal(f1, f)
    assert np.all(dip[:-1, 2:] == 0) and np.all(tm[:-1, 3:] == 0) and np.any(dip[-1, 2:] != 0)


def test_cbt_dipoles_and_transfer_matrices():
    """Geometry of the reconstructed dipole: the two charges sit delta = 0.3 nm apart, centred on the middle of the
    b-c bond.  The transfer matrices: with the sign of the three gamma-derivative terms flipped (see
    ``bonded_oracle.reconstruct``) they are the exact Jacobians d d / d r_i at fixed phi -- this pins every other
    term of the restatement; the default keeps the reference's sign."""
    rng = np.random.default_rng(3)
    ra, rb, rc = rng.normal(size=(3, 3)) + 2.0
    box = np.array([50.0, 50.0, 50.0])
    c_k, d_k, phi = np.array([30.0, 5, 2, 1, 0.5]), np.array([0.1, 0.2, -0.3, 0.4, 0.0]), 0.7

    def half(ra, rb, rc, sign):
        rec = bo.reconstruct(ra - rb, rb, rc - rb, box, c_k, d_k, phi, 1, gamma_sign=sign)
        return 0.5 * (rec[5][0] - rec[5][1]), 0.5 * (rec[5][0] + rec[5][1]), rec[6]
    d, centre, T_ref = half(ra, rb, rc, 1.0)
    assert np.linalg.norm(2 * d) == pytest.approx(0.3, rel=1e-6)     # cos_psi, sin_psi are single-precision constants
    np.testing.assert_allclose(centre, 0.5 * (rb + rc), atol=1e-12)
    _, _, T = half(ra, rb, rc, -1.0)
    assert np.abs(T - T_ref).max() > 1e-2 * np.abs(T).max()
    h = 1e-6
    for which in range(3):
        J = np.zeros((3, 3))
        for k in range(3):
            p, m = [ra.copy(), rb.copy(), rc.copy()], [ra.copy(), rb.copy(), rc.copy()]
            p[which][k] += h
            m[which][k] -= h
            J[k] = (half(*p, -1.0)[0] - half(*m, -1.0)[0]) / (2 * h)
        np.testing.assert_allclose(T[which], J, rtol=0, atol=1e-8)


def test_dipole_force_redistribution_is_the_chain_rule():
    """hymd/force.py:855-880: f_bead = sum_i D_i (f_+ - f_-) + 1/2 (f_+ + f_-) on the two beads of the bond that
    carries the dipole; total force is conserved for the centre term."""
    r, box, idx, coeff, dt, last = _cbt_system(seed=5)
    _, _, dip, tm = bo.compute_dihedral_forces(r, box, *idx, coeff, dt, last, dipole_flag=1, full=True)
    rng = np.random.default_rng(2)
    fd = rng.normal(size=dip.shape)
    out = bo.dipole_forces_redistribution(len(r), fd, tm, *idx, dt, last)
    want = np.zeros_like(out)
    for t, (i, j, k, l) in enumerate(zip(*idx)):
        s, df = fd[t, 0] + fd[t, 1], fd[t, 0] - fd[t, 1]
        want[i] += tm[t, 0].astype(float) @ df
        want[j] += tm[t, 1].astype(float) @ df + 0.5 * s
        want[k] += tm[t, 2].astype(float) @ df + 0.5 * s
        if last[t]:
            s, df = fd[t, 2] + fd[t, 3], fd[t, 2] - fd[t, 3]
            want[j] += tm[t, 3].astype(float) @ df
            want[k] += tm[t, 4].astype(float) @ df + 0.5 * s
            want[l] += tm[t, 5].astype(float) @ df + 0.5 * s
    np.testing.assert_allclose(out, want, atol=1e-12)

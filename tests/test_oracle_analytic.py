"""Anchor the quantities the reference leaves unpinned (forces, -ik gradient, PME) on
analytic results; see oracle/__init__.py "Pinning status"."""
import numpy as np
import pytest

from oracle import field_oracle as fo
from oracle import pm_oracle as pmo
from oracle.analytic import (ewald_reciprocal_on_mesh_k, gaussian_core_energy,
                             gaussian_core_forces)
from oracle.hamiltonian_oracle import OracleHamiltonian
from conftest import make_config

R5 = np.array([[1.50, 0.75, 2.25], [2.25, 0.00, 3.00], [4.50, 1.50, 2.25],
               [0.75, 3.00, 0.75], [3.00, 2.25, 1.50]])


def test_field_forces_match_gaussian_core_pair_forces():
    """Mesh forces (paint -> filter -> v_ext -> filter -> -ik -> readout) equal the analytic
    Gaussian-core pair forces (hymd/gaussian_core.py:34-54) for on-vertex particles."""
    mesh = [160, 160, 160]
    chi = [("A", "B", 9.6754032616815161), ("A", "C", -13.2596290315913623),
           ("B", "C", 0.3852001771213374)]
    names = ["A", "B", "C", "A", "B"]
    tid = np.array([0, 1, 2, 0, 1])
    kappa, sigma = 0.05, 0.5
    cfg = make_config(names, 5, mesh, [15.0, 15.0, 15.0], sigma=sigma, kappa=kappa, chi=chi)
    W = OracleHamiltonian(cfg)
    st = fo.FieldState(cfg, np.float64)
    fo.update_field(st, W, R5, tid, cfg, workers=-1)
    f = fo.compute_field_force(st, R5, tid, 3)
    ref = gaussian_core_forces(R5, tid, W.chi, kappa, sigma, cfg.rho0)
    assert np.abs(f - ref).max() / np.abs(ref).max() < 1e-6
    e, _, _ = fo.compute_field_and_kinetic_energy(st, W, np.zeros((5, 3)), cfg)
    assert e == pytest.approx(gaussian_core_energy(R5, tid, W.chi, kappa, sigma, cfg.rho0), abs=1e-5)


def test_pme_matches_gaussian_smeared_ewald_sum():
    mesh = [25, 25, 25]
    box = [5.0, 5.0, 5.0]
    rng = np.random.default_rng(11)
    idx = rng.integers(0, 25, size=(12, 3))
    r = idx * (5.0 / 25)
    q = rng.normal(size=12)
    q -= q.mean()
    cfg = make_config(["A"], 12, mesh, box, sigma=0.35, hamiltonian="DefaultNoChi",
                      coulombtype="PIC_Spectral", dielectric_const=80.0)
    W = OracleHamiltonian(cfg)
    st = fo.FieldState(cfg, np.float64)
    fo.update_field(st, W, r, np.zeros(12, dtype=int), cfg)
    f = fo.update_field_force_q(st, W, q, r, cfg)
    conv = cfg.coulomb_constant / cfg.dielectric_const
    e_ref, f_ref = ewald_reciprocal_on_mesh_k(q, r, mesh, cfg.box_size, cfg.sigma, conv)
    cfg.self_energy = fo.compute_self_energy_q(cfg, q)
    W = OracleHamiltonian(cfg)
    _, _, e_q = fo.compute_field_and_kinetic_energy(st, W, np.zeros((12, 3)), cfg)
    assert e_q + cfg.self_energy == pytest.approx(e_ref, rel=1e-12)
    assert np.abs(f - f_ref).max() / np.abs(f_ref).max() < 1e-12


@pytest.mark.parametrize("mesh", [(16, 12, 10), (8, 8, 8), (9, 12, 10), (6, 5, 7), (24, 24, 24)])
def test_nyquist_minimal_zeroing_rule(mesh):
    """-i k_d on the stored half spectrum (fftfreq sign at Nyquist) followed by irfftn equals the
    Hermitian-consistent spectrum obtained by the minimal zeroing rule of SURVEY.md section 7:
    d=z: zero the k_z = N_z/2 plane; d in {x,y}: zero the index_d = N_d/2 line inside the
    self-conjugate planes k_z in {0, N_z/2}.  This is what the fused k-space kernel feeds cuFFT."""
    rng = np.random.default_rng(3)
    box = np.array([3.0, 2.5, 4.0])
    x = rng.normal(size=mesh)
    vf = pmo.r2c(x)
    k = pmo.kgrid(mesh, box)
    nx, ny, nz = mesh
    for d in range(3):
        raw = pmo.c2r(-1j * k[d] * vf, mesh)
        keff = [np.broadcast_to(kk, vf.shape).copy() for kk in k]
        selfconj = [0] + ([nz // 2] if nz % 2 == 0 else [])
        if d == 2:
            if nz % 2 == 0:
                keff[2][:, :, nz // 2] = 0.0
        elif d == 0:
            if nx % 2 == 0:
                for kz in selfconj:
                    keff[0][nx // 2, :, kz] = 0.0
        else:
            if ny % 2 == 0:
                for kz in selfconj:
                    keff[1][:, ny // 2, kz] = 0.0
        ruled = pmo.c2r(-1j * keff[d] * vf, mesh)
        assert np.abs(raw - ruled).max() <= 1e-13 * max(1.0, np.abs(raw).max())
        # and the ruled half spectrum is exactly the half of a Hermitian full spectrum
        full = np.fft.fftn(ruled) / ruled.size
        np.testing.assert_allclose(full[:, :, : nz // 2 + 1], -1j * keff[d] * vf, atol=1e-12)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_c_and_numpy_cic_agree(dtype):
    rng = np.random.default_rng(2)
    box = np.array([4.0, 5.0, 6.0])
    mesh = (12, 10, 9)
    r = rng.uniform(0, 1, size=(2000, 3)) * box
    r[:3] = [[0, 0, 0], [box[0] - 1e-9, box[1] - 1e-9, box[2] - 1e-9], [2.0, 2.5, 3.0]]
    m = rng.uniform(0.5, 2.0, size=2000)
    a = pmo.cic_paint(r, m, mesh, box, dtype, use_c=False)
    b = pmo.cic_paint(r, m, mesh, box, dtype, use_c=True)
    tol = 1e-12 if dtype == np.float64 else 2e-5
    assert np.abs(a - b).max() < tol
    assert a.sum(dtype=np.float64) == pytest.approx(m.astype(dtype).sum(dtype=np.float64), rel=1e-6)
    g = rng.normal(size=mesh).astype(dtype)
    va = pmo.cic_readout(g, r, box, use_c=False)
    vb = pmo.cic_readout(g, r, box, use_c=True)
    assert np.abs(va - vb).max() < tol
    # multi-threaded variants (CPU baseline) give the same numbers
    c = fo._paint_mt(r, m, mesh, box, dtype)
    assert np.abs(c - b).max() < tol
    assert np.abs(fo._readout_mt(g, r, box) - vb).max() == 0.0


def test_laplacian_of_a_plane_wave():
    """comp_laplacian (field.py:406-425) on phi = cos(q.x): every component is -q_d^2 cos(q.x)."""
    from types import SimpleNamespace
    from oracle import field_oracle as fo
    from oracle import pm_oracle as pmo
    mesh, box = (16, 12, 10), np.array([4.0, 5.0, 6.0])
    n = (2, 1, 3)
    q = 2.0 * np.pi * np.array(n) / box
    x = [np.arange(mesh[a]) * box[a] / mesh[a] for a in range(3)]
    phase = q[0] * x[0][:, None, None] + q[1] * x[1][None, :, None] + q[2] * x[2][None, None, :]
    phi = np.cos(phase)
    cfg = SimpleNamespace(mesh_size=list(mesh), box_size=box, n_types=1)
    st = fo.FieldState(SimpleNamespace(mesh_size=list(mesh), box_size=box, n_types=1, dtype=np.float64))
    st.phi_fourier[0] = pmo.r2c(phi)
    lap = fo.comp_laplacian(st, cfg)
    for d in range(3):
        np.testing.assert_allclose(lap[0][d], -q[d] ** 2 * phi, atol=1e-12 * q[d] ** 2 + 1e-13)

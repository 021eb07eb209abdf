"""Pin the oracle against the reference's own known-answer tests for this path.

Each test re-expresses a reference test with the oracle standing in for pmesh:
* ``test/test_hamiltonian.py:23-52``   filter value
* ``test/test_hamiltonian.py:62-128``  mass conservation of paint and filter
* ``test/test_hamiltonian.py:172-227`` 3 particles / 1 type Gaussian-core energy (abs 1e-6)
* ``test/test_hamiltonian.py:270-360`` 5 particles / 3 types with chi (abs 1e-6)
"""
import numpy as np
import pytest

from oracle import pm_oracle as pmo
from oracle.analytic import gaussian_core_energy
from oracle.field_oracle import volume_per_cell
from oracle.hamiltonian_oracle import OracleHamiltonian
from conftest import make_config


@pytest.mark.parametrize("sigma", [0.2988365823859701, 1.2585762493242553, 9.2159828579248931])
def test_window_function(sigma):
    cfg = make_config(["A"], 3, 8, [5.0, 5.0, 5.0], sigma=sigma, hamiltonian="DefaultNoChi")
    W = OracleHamiltonian(cfg)
    k_ = np.array([[0.5321315378106508, -0.6711591309063634, 0.8362051282174443],
                   [0.2853046917286570, -0.6542962742862817, 0.3174805390299977],
                   [0.6999748102259762, -0.9385345654631219, -0.7383831543700541]])
    v = np.array([[0.6106860760556785, -0.5406324770662296, 0.6388756736156205],
                  [-0.7348831910188103, 0.2808258965802970, 0.7446817693106476],
                  [0.6458163432308923, -0.0526126093343278, 0.7065510160449484]])
    for kk, vv in zip(k_, v):
        expect = vv * np.exp(-0.5 * sigma ** 2 * np.dot(kk, kk))
        assert np.allclose(W.H(kk, vv), expect, atol=1e-14)


@pytest.mark.parametrize("use_c", [False, True])
def test_paint_and_filter_conserve_mass(use_c):
    # test_hamiltonian.py:62-128: 8^3 mesh, box [7.1598, 11.2498, 5.1009]
    box = np.array([7.1598, 11.2498, 5.1009])
    mesh = (8, 8, 8)
    rng = np.random.default_rng(5)
    r = rng.uniform(0, 1, size=(5, 3)) * box
    cfg = make_config(["A", "B"], 5, list(mesh), box, sigma=0.2988365823859701,
                      kappa=0.029230985982, hamiltonian="DefaultNoChi")
    dv = volume_per_cell(cfg)
    phi = pmo.cic_paint(r, 1.0, mesh, cfg.box_size, np.float64, use_c) / dv
    assert pmo.csum(phi) == pytest.approx(5 / dv, abs=1e-12)
    W = OracleHamiltonian(cfg)
    k = pmo.kgrid(mesh, cfg.box_size)
    phit = pmo.c2r(W.H(k, pmo.r2c(phi)), mesh)
    assert pmo.csum(phit) == pytest.approx(5 / dv, abs=1e-10)


R3 = np.array([[1.50, 0.75, 2.25], [2.25, 0.00, 3.00], [4.50, 1.50, 2.25]])


@pytest.mark.parametrize("kappa,sigma", [(0.029230985982, 0.2988365823859701),
                                         (1.299759825895, 1.2095870248085025)])
def test_no_chi_gaussian_core_energy(kappa, sigma):
    mesh = (160, 160, 160)
    cfg = make_config(["A"], 3, list(mesh), [15.0, 15.0, 15.0], sigma=sigma, kappa=kappa,
                      hamiltonian="DefaultNoChi")
    dv = volume_per_cell(cfg)
    phi = pmo.cic_paint(R3, 1.0, mesh, cfg.box_size, np.float64) / dv
    assert pmo.csum(phi) == pytest.approx(3 / dv, abs=1e-9)
    k = pmo.kgrid(mesh, cfg.box_size)
    for kind, shift in (("SquaredPhi", False), ("DefaultNoChi", True)):
        W = OracleHamiltonian(cfg, kind)
        phit = pmo.c2r(W.H(k, pmo.r2c(phi, workers=-1)), mesh, workers=-1)
        w = pmo.csum(W.w_0([phit]) * dv)
        E = gaussian_core_energy(R3, [0, 0, 0], np.zeros((1, 1)), kappa, sigma, cfg.rho0, shift)
        assert w == pytest.approx(E, abs=1e-6)


R5 = np.array([[1.50, 0.75, 2.25], [2.25, 0.00, 3.00], [4.50, 1.50, 2.25],
               [0.75, 3.00, 0.75], [3.00, 2.25, 1.50]])


@pytest.mark.parametrize("kappa,sigma", [(0.029230985982, 0.2988365823859701),
                                         (1.299759825895, 1.2095870248085025)])
def test_with_chi_gaussian_core_energy(kappa, sigma):
    # structure of test_hamiltonian.py:270-360 (5 particles, 3 types, the reference's chi values)
    mesh = (160, 160, 160)
    chi = [("A", "B", 9.6754032616815161), ("A", "C", -13.2596290315913623),
           ("B", "C", 0.3852001771213374)]
    names = ["A", "B", "C", "A", "B"]
    tid = np.array([0, 1, 2, 0, 1])
    cfg = make_config(names, 5, list(mesh), [15.0, 15.0, 15.0], sigma=sigma, kappa=kappa,
                      hamiltonian="DefaultWithChi", chi=chi)
    dv = volume_per_cell(cfg)
    W = OracleHamiltonian(cfg)
    k = pmo.kgrid(mesh, cfg.box_size)
    phit = []
    for t in range(3):
        phi = pmo.cic_paint(R5[tid == t], 1.0, mesh, cfg.box_size, np.float64) / dv
        phit.append(pmo.c2r(W.H(k, pmo.r2c(phi, workers=-1)), mesh, workers=-1))
    w = pmo.csum(W.w_0(phit) * dv)
    E = gaussian_core_energy(R5, tid, W.chi, kappa, sigma, cfg.rho0, True)
    assert w == pytest.approx(E, abs=1e-6)

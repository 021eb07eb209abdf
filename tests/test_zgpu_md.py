"""GPU parity of row f2 (bonded forces, rRESPA updates, CSVR thermostat) against the oracle and the
reference's golden vectors, through hymd_b200.force / .thermostat / .md (ctypes -> C ABI).

Tolerances: forces 1e-10 (fp64) / 1e-6 (fp32: one rounding of the float64 force to float32; the oracle
takes position differences in float32 like the Fortran), relative to the largest force; energies and
pressure by-products 1e-11 / 1e-6.  The file name sorts after the field-force parity tests on purpose:
these kernels were added after the round's GPU budget was spent."""
import os

import numpy as np
import pytest
import torch

from oracle import bonded_oracle as bo
from oracle import thermostat_oracle as to

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "bonded_golden.npz"))
TG = np.load(os.path.join(HERE, "golden", "thermostat_golden.npz"))
FTOL = {np.float32: 1e-6, np.float64: 1e-10}
ETOL = {np.float32: 1e-6, np.float64: 1e-11}


def chains(rng, n_chains, length, box, real):
    r = np.empty((n_chains * length, 3))
    for m in range(n_chains):
        steps = rng.normal(size=(length - 1, 3))
        steps *= 0.47 / np.linalg.norm(steps, axis=1)[:, None]
        r[m * length:(m + 1) * length] = rng.random(3) * box + np.concatenate([np.zeros((1, 3)), np.cumsum(steps, 0)])
    r = np.mod(r, box).astype(real)
    first = (np.arange(n_chains) * length)[:, None]
    a2 = (first + np.arange(length - 1)[None, :]).ravel()
    a3 = (first + np.arange(length - 2)[None, :]).ravel()
    a4 = (first + np.arange(length - 3)[None, :]).ravel()
    return r, a2, a3, a4


DEVICE = "cuda"       # tests/test_md_host_emulation.py re-runs these bodies on "cpu" through a host shim


def dev(x, real):
    return torch.tensor(np.ascontiguousarray(x), dtype=torch.float64 if real == np.float64 else torch.float32,
                        device=DEVICE)        # always a copy (as_tensor would alias x on the CPU dry run)


@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_bonded_forces_match_oracle(real):
    from hymd_b200 import force as F
    rng = np.random.default_rng(31)
    box = np.array([5.0, 6.0, 7.0])
    r, a2, a3, a4 = chains(rng, 300, 11, box, real)     # 3300 particles: several CTAs, ragged tail
    n = len(r)
    r0, k2 = 0.47 + 0.1 * rng.random(len(a2)), 1250.0 * (0.5 + rng.random(len(a2)))
    t0, k3 = np.radians(rng.choice([120.0, 180.0], size=len(a3))), 25.0 * (0.5 + rng.random(len(a3)))
    coeff = np.zeros((len(a4), 6, 5))
    coeff[:, 0] = rng.normal(size=(len(a4), 5)) * 4
    coeff[:, 1] = rng.uniform(-np.pi, np.pi, size=(len(a4), 5))
    coeff[1::3, 2] = rng.normal(size=coeff[1::3, 2].shape)
    coeff[1::3, 3] = rng.normal(size=coeff[1::3, 3].shape)
    dt4 = np.zeros(len(a4), dtype=np.int64)
    dt4[::3] = 2
    coeff[::3, 0, 0], coeff[::3, 0, 1] = 0.3, 40.0
    pos = dev(r, real)

    f = torch.full((n, 3), 7.0, dtype=pos.dtype, device=DEVICE)      # must be overwritten
    e, pr = F.compute_bond_forces(f, pos, box, a2, a2 + 1, r0, k2)
    fo_, eo, pro = bo.compute_bond_forces(r, box, a2, a2 + 1, r0, k2)
    assert np.abs(f.cpu().numpy() - fo_).max() <= FTOL[real] * np.abs(fo_).max()
    assert float(e) == pytest.approx(eo, rel=ETOL[real])
    np.testing.assert_allclose(pr.cpu().numpy(), pro, rtol=0, atol=ETOL[real] * np.abs(pro).max())

    f = torch.full((n, 3), 7.0, dtype=pos.dtype, device=DEVICE)
    e, pr = F.compute_angle_forces(f, pos, box, a3, a3 + 1, a3 + 2, t0, k3)
    fo_, eo, pro = bo.compute_angle_forces(r, box, a3, a3 + 1, a3 + 2, t0, k3)
    assert np.abs(f.cpu().numpy() - fo_).max() <= FTOL[real] * np.abs(fo_).max()
    assert float(e) == pytest.approx(eo, rel=ETOL[real])
    np.testing.assert_allclose(pr.cpu().numpy(), pro, rtol=0, atol=ETOL[real] * np.abs(pro).max())

    f = torch.full((n, 3), 7.0, dtype=pos.dtype, device=DEVICE)
    e = F.compute_dihedral_forces(f, pos, None, None, box, a4, a4 + 1, a4 + 2, a4 + 3, coeff, dt4)
    fo_, eo = bo.compute_dihedral_forces(r, box, a4, a4 + 1, a4 + 2, a4 + 3, coeff, dt4)
    assert np.abs(f.cpu().numpy() - fo_).max() <= FTOL[real] * np.abs(fo_).max()
    assert float(e) == pytest.approx(eo, rel=ETOL[real])


@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_cbt_dihedrals_dipoles_and_redistribution_match_oracle(real):
    """dih_type 1 through the reference-shaped interface (``cdf`` with ``bb_index`` / ``dipole_flag``, then
    ``dipole_forces_redistribution``): forces, energy, dipole positions, transfer matrices and redistributed forces
    against the oracle's restatement of compute_dihedral_forces.f90 + dipole_reconstruction.f90 + force.py:855-880;
    device tensors and the numpy interface; mixed with dihedrals of types 0 and 2."""
    from hymd_b200 import _lib
    from hymd_b200 import force as F
    rng = np.random.default_rng(33)
    n = 300
    box = np.array([3.0, 3.5, 4.0])
    r = np.mod(np.cumsum(rng.normal(scale=0.33, size=(n, 3)), axis=0) + 1.5, box).astype(real)
    a = np.arange(n - 3)
    nt = len(a)
    coeff = np.zeros((nt, 6, 5))
    coeff[:, 0] = rng.uniform(0.5, 2.0, size=(nt, 5))
    coeff[:, 1] = rng.uniform(-1, 1, size=(nt, 5))
    coeff[:, 4] = rng.uniform(20, 60, size=(nt, 5))
    coeff[:, 5] = rng.uniform(-1, 1, size=(nt, 5))
    dt = rng.choice([0, 1, 1, 2], size=nt)
    dt[-1] = 1
    coeff[dt == 2, 0, :2] = [0.4, 30.0]
    last = np.zeros(nt, dtype=int)
    last[-1] = 1
    last[np.nonzero(dt == 1)[0][::7]] = 1
    idx = (a, a + 1, a + 2, a + 3)
    fo_, eo, dipo, tmo = bo.compute_dihedral_forces(r, box, *idx, coeff, dt, last, dipole_flag=1, full=True)
    tol = 1e-10 if real == np.float64 else 1e-6
    scale = np.abs(fo_).max()
    tdt = torch.float64 if real == np.float64 else torch.float32
    # device tensors
    pos = torch.as_tensor(r, device=DEVICE)
    f = torch.empty_like(pos)
    dip = torch.zeros((nt, 4, 3), dtype=tdt, device=DEVICE)
    tm = torch.zeros((nt, 6, 3, 3), dtype=tdt, device=DEVICE)
    e = F.compute_dihedral_forces(f, pos, dip, tm, box, *idx, coeff, dt, last, 1)
    assert np.abs(f.cpu().numpy() - fo_).max() < tol * scale
    assert float(e) == pytest.approx(eo, rel=tol)
    assert np.abs(dip.cpu().numpy().astype(np.float64) - dipo.astype(np.float64)).max() < (1e-12 if real == np.float64 else 2e-6)
    assert np.abs(tm.cpu().numpy().astype(np.float64) - tmo.astype(np.float64)).max() < tol * max(1.0, np.abs(tmo).max())
    # dipole_flag = 0: same forces and energy, dipole arrays zeroed (compute_dihedral_forces.f90:27-28)
    f0 = torch.empty_like(pos)
    e0 = F.compute_dihedral_forces(f0, pos, dip, tm, box, *idx, coeff, dt, last, 0)
    assert torch.equal(f0, f) and float(e0) == float(e) and not dip.any() and not tm.any()
    # redistribution of forces on the dipole charges (what the second PME call returns) to the beads
    fd = rng.normal(size=(nt, 4, 3)).astype(real)
    want = bo.dipole_forces_redistribution(n, fd, tmo, *idx, dt, last)
    fb = torch.full((n, 3), 7.0, dtype=tdt, device=DEVICE)
    F.dipole_forces_redistribution(fb, torch.as_tensor(fd, device=DEVICE), torch.as_tensor(tmo, device=DEVICE),
                                   *idx, dt, last, coeff=coeff)
    assert np.abs(fb.cpu().numpy() - want).max() < tol * max(1.0, np.abs(want).max())
    # numpy interface, Fortran-ordered arrays as main.py:505-506 makes them
    fn = np.zeros((n, 3), dtype=real)
    dipn = np.asfortranarray(np.ones((nt, 4, 3), dtype=real))
    tmn = np.asfortranarray(np.ones((nt, 6, 3, 3), dtype=real))
    en = F.compute_dihedral_forces(fn, r, dipn, tmn, box, *idx, coeff, dt, last, 1)
    assert np.abs(fn - fo_).max() < tol * scale and en == pytest.approx(eo, rel=tol)
    assert np.abs(dipn.astype(np.float64) - dipo.astype(np.float64)).max() < (1e-12 if real == np.float64 else 2e-6)
    fbn = np.zeros((n, 3), dtype=real)
    F.dipole_forces_redistribution(fbn, fd, tmn, *idx, dt, last)
    assert np.abs(fbn - want).max() < 2 * tol * max(1.0, np.abs(want).max())
    # the fused inner step carries the bending term: the dihedral force array it hands out is cdf's
    topo = F.BondedTopology(n, dihedrals=(*idx, coeff, dt, last), device=DEVICE)
    f3 = [None, None, torch.zeros_like(pos)]
    res = topo.inner_step(pos, None, torch.zeros_like(pos), box, 72.0, 0.01, 1, 0.0, force_out=f3)
    assert np.abs(f3[2].cpu().numpy() - fo_).max() < tol * scale
    assert float(res[2][0]) == pytest.approx(eo, rel=tol)


def test_reference_kats_on_the_device_and_numpy_interface():
    """test/test_force.py known answers (via the golden file) with numpy in / numpy out like the f2py
    kernels; the dihedral sign is the production Fortran's (see tests/test_oracle_bonded.py)."""
    from hymd_b200 import force as F
    r, box = G["dppc/r"], G["dppc/box"]
    a, b, r0, k = (G["dppc/b2_" + x] for x in ("a", "b", "r0", "k"))
    f = np.zeros_like(r)
    e, pr = F.compute_bond_forces(f, r, box, a, b, r0, k)
    assert isinstance(e, float) and isinstance(pr, np.ndarray)
    np.testing.assert_allclose(f, G["dppc/b2_force"], rtol=0, atol=1e-10)
    assert e == pytest.approx(float(G["dppc/b2_energy"]), abs=1e-11)
    a, b, c, t0, k = (G["dppc/b3_" + x] for x in ("a", "b", "c", "t0", "k"))
    f = np.asfortranarray(np.zeros_like(r))           # main.py:500-506 hands Fortran-ordered arrays
    e, pr = F.compute_angle_forces(f, np.asfortranarray(r), box, a, b, c, t0, k)
    np.testing.assert_allclose(f, G["dppc/b3_force"], rtol=0, atol=1e-9)
    assert e == pytest.approx(float(G["dppc/b3_energy"]), abs=1e-10)
    r, box = G["ala/r"], G["ala/box"]
    a, b, c, d, coeff, dt = (G["ala/" + x] for x in ("a", "b", "c", "d", "coeff", "dtype"))
    f = np.zeros_like(r)
    dip, tm = np.ones((5, 4, 3)), np.ones((5, 6, 3, 3))
    e = F.compute_dihedral_forces(f, r, dip, tm, box, a, b, c, d, coeff, dt, np.zeros(5, dtype=int), 0)
    np.testing.assert_allclose(f, -G["ala/term_force_plain"].sum(axis=0), rtol=0, atol=1e-10)
    assert e == pytest.approx(float(G["ala/term_energy"].sum()), abs=1e-11)
    assert not dip.any() and not tm.any()


def test_bonded_edge_cases():
    from hymd_b200 import _lib
    from hymd_b200 import force as F
    box = np.array([3.0, 3.0, 3.0])
    pos = torch.rand((5, 3), dtype=torch.float64, device=DEVICE)
    f = torch.ones_like(pos)
    empty = np.zeros(0, dtype=np.int64)
    e, pr = F.compute_bond_forces(f, pos, box, empty, empty, np.zeros(0), np.zeros(0))   # no terms: f = 0
    assert float(e) == 0.0 and not f.any()
    with pytest.raises(_lib.HymdError):          # index outside the local particles
        F.compute_bond_forces(f, pos, box, np.array([0]), np.array([9]), np.array([0.4]), np.array([1.0]))
    with pytest.raises(_lib.HymdError):          # dih_type outside 0, 1, 2
        F.compute_dihedral_forces(f, pos, None, None, box, np.array([0]), np.array([1]), np.array([2]),
                                  np.array([3]), np.zeros((1, 6, 5)), np.array([3]))
    # bitwise reproducible
    a = np.arange(4)
    args = (box, a, a + 1, np.full(4, 0.4), np.full(4, 100.0))
    f1, f2 = torch.empty_like(pos), torch.empty_like(pos)
    e1, _ = F.compute_bond_forces(f1, pos, *args)
    e2, _ = F.compute_bond_forces(f2, pos, *args)
    assert torch.equal(f1, f2) and float(e1) == float(e2)


@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_kick_drift_matches_numpy(real):
    from hymd_b200.md import kick_drift
    rng = np.random.default_rng(5)
    n = 70001
    box = np.array([3.0, 4.0, 5.0])
    x = (rng.random((n, 3)) * box).astype(real)
    v = rng.normal(scale=2.0, size=(n, 3)).astype(real)
    fs = [rng.normal(scale=300.0, size=(n, 3)).astype(real) for _ in range(3)]
    mass, dt = 72.0, 0.03
    xd, vd, fd = dev(x, real), dev(v, real), [dev(f, real) for f in fs]
    kick_drift(vd, xd, fd, mass, dt, dt, box)
    v_ref = v + real(0.5 * dt) * ((fs[0] + fs[1] + fs[2]) / real(mass))
    x_ref = np.mod(x + real(dt) * v_ref, box.astype(real)[None, :])
    eps = np.finfo(real).eps
    assert np.abs(vd.cpu().numpy() - v_ref).max() <= 2 * eps * np.abs(v_ref).max()   # FMA contraction only
    xg = xd.cpu().numpy()
    wrapdiff = np.abs(xg - x_ref)
    wrapdiff = np.minimum(wrapdiff, np.abs(wrapdiff - box.astype(real)[None, :]))
    assert wrapdiff.max() <= 8 * eps * box.max()
    assert (xg >= 0).all() and (xg < box.astype(real)).all()
    vd2 = dev(v, real)
    kick_drift(vd2, None, fd[:2], mass, 5 * dt, sequential=True)
    v_ref = v + real(0.5 * 5 * dt) * (fs[0] / real(mass))
    v_ref = v_ref + real(0.5 * 5 * dt) * (fs[1] / real(mass))
    assert np.abs(vd2.cpu().numpy() - v_ref).max() <= 2 * eps * np.abs(v_ref).max()


class _Cfg:
    gas_constant = 0.0083144621

    def __init__(self, names, groups, params, n_particles):
        self.mass, _, self.target_temperature, self.time_step, inner, self.tau = [float(x) for x in params]
        self.respa_inner = int(inner)
        self.gas_constant = float(params[1])
        self.unique_names = sorted({n.decode() for n in names})
        self.name_to_type_map = {n: i for i, n in enumerate(self.unique_names)}
        self.thermostat_coupling_groups = [list(g) for g in groups]
        self.thermostat_work = 0.0
        self.n_particles = n_particles


class _Mock:
    def __init__(self, x):
        self.x, self.i = list(x), 0

    def __call__(self, *args):
        self.i += 1
        return self.x[self.i - 1]


@pytest.mark.parametrize("remove", [False, True])
@pytest.mark.parametrize("case,groups", [("all", []), ("abcd", [["A"], ["B"], ["C"], ["D"]]),
                                         ("abc_d", [["A", "B", "C"], ["D"]])])
def test_csvr_matches_reference_golden(case, groups, remove):
    """The reference's own csvr_thermostat outputs (test/test_thermostat.py fixture and draws)."""
    from hymd_b200 import thermostat as T
    pre = f"mws/{case}_{'com' if remove else 'nocom'}"
    names = TG["mws/names"]
    cfg = _Cfg(names, groups, TG["mws/params"], len(names))
    v = dev(TG["mws/velocities"], np.float64)
    T.csvr_thermostat(v, names, cfg, None, random_gaussian=_Mock(TG[pre + "/gauss"]),
                      random_chi_squared=_Mock(TG[pre + "/chi2"]), remove_center_of_mass_momentum=remove)
    np.testing.assert_allclose(v.cpu().numpy(), TG[pre + "/v"], rtol=1e-11, atol=1e-13)
    assert float(cfg.thermostat_work) == pytest.approx(float(TG[pre + "/work"]), abs=1e-10)


def test_csvr_larger_system_numpy_interface_and_cancel_com():
    from hymd_b200 import thermostat as T
    names = TG["rand/names"]
    cfg = _Cfg(names, [["A", "B"], ["W"]], TG["rand/params"], len(names))
    v = TG["rand/v0"].copy()
    T.csvr_thermostat(v, names, cfg, None, random_gaussian=_Mock(TG["rand/gauss"]),
                      random_chi_squared=_Mock(TG["rand/chi2"]))
    np.testing.assert_allclose(v, TG["rand/v"], rtol=1e-11, atol=1e-13)
    assert isinstance(cfg.thermostat_work, float)
    assert cfg.thermostat_work == pytest.approx(float(TG["rand/work"]), rel=1e-10)
    # float32 velocities, type ids instead of names
    cfg = _Cfg(names, [["A", "B"], ["W"]], TG["rand/params"], len(names))
    types = torch.as_tensor(np.array([cfg.name_to_type_map[n.decode()] for n in names], dtype=np.int32), device=DEVICE)
    v32 = dev(TG["rand/v0"], np.float32)
    T.csvr_thermostat(v32, types, cfg, None, random_gaussian=_Mock(TG["rand/gauss"]),
                      random_chi_squared=_Mock(TG["rand/chi2"]))
    ref = TG["rand/v0"].astype(np.float32).astype(np.float64)
    grp = np.where(names == b"W", 1, 0).astype(np.int32)
    mass, gas, T0, dt, inner, tau = TG["rand/params"]
    to.csvr_thermostat(ref, grp, 2, mass=mass, gas_constant=gas, target_temperature=T0, time_step=dt,
                       respa_inner=int(inner), tau=tau, draws=list(zip(TG["rand/gauss"], TG["rand/chi2"])))
    assert np.abs(v32.cpu().numpy() - ref).max() <= 2e-7 * np.abs(ref).max()
    # cancel_com_momentum (thermostat.py:12-15)
    cfg = _Cfg(TG["mws/names"], [], TG["mws/params"], len(TG["mws/names"]))
    v = dev(TG["mws/velocities"], np.float64)
    T.cancel_com_momentum(v, cfg)
    np.testing.assert_allclose(v.cpu().numpy(), TG["mws/cancel_com"], rtol=1e-13, atol=1e-15)
    assert float(T.kinetic_energy(dev(TG["mws/velocities"], np.float64), 72.0)) == pytest.approx(
        168.45555165866017, abs=1e-11)                     # test_thermostat.py:74


@pytest.mark.parametrize("fused", [True, False])
def test_respa_md_with_bonds_conserves_energy_like_the_oracle(fused):
    """A short NVE run of bonded chains (no field forces: the slow-force callback returns nothing):
    the device rRESPA step and a float64 numpy velocity-Verlet driven by the bonded oracle follow the
    same trajectory, and bonded + kinetic energy is conserved."""
    from hymd_b200.force import BondedTopology
    from hymd_b200.md import RespaMD
    rng = np.random.default_rng(9)
    box = np.array([4.0, 4.0, 4.0])
    r, a2, a3, _ = chains(rng, 40, 8, box, np.float64)
    n = len(r)
    r0, k2 = np.full(len(a2), 0.47), np.full(len(a2), 1250.0)
    t0, k3 = np.full(len(a3), np.radians(120.0)), np.full(len(a3), 25.0)
    v = rng.normal(scale=0.15, size=(n, 3))
    mass, dt, inner, steps = 72.0, 0.005, 4, 25

    def oracle_forces(x):
        fb, eb, _ = bo.compute_bond_forces(x, box, a2, a2 + 1, r0, k2)
        fa, ea, _ = bo.compute_angle_forces(x, box, a3, a3 + 1, a3 + 2, t0, k3)
        return fb + fa, eb + ea
    xo, vo = r.copy(), v.copy()
    fo_, e_pot = oracle_forces(xo)
    e0 = e_pot + 0.5 * mass * np.sum(vo ** 2)
    for _ in range(steps * inner):
        vo = vo + 0.5 * dt * (fo_ / mass)
        xo = np.mod(xo + dt * vo, box[None, :])
        fo_, e_pot = oracle_forces(xo)
        vo = vo + 0.5 * dt * (fo_ / mass)
    e1 = e_pot + 0.5 * mass * np.sum(vo ** 2)
    assert abs(e1 - e0) < 2e-3 * abs(e0)

    topo = BondedTopology(n, bonds=(a2, a2 + 1, r0, k2), angles=(a3, a3 + 1, a3 + 2, t0, k3))
    fout = [torch.zeros((n, 3), dtype=torch.float64, device=DEVICE) for _ in range(2)] + [None]
    md = RespaMD(lambda x: [], box, mass, dt, respa_inner=inner, topology=topo, fused=fused,
                 force_out=fout if fused else None)
    xd, vd = dev(r, np.float64), dev(v, np.float64)
    slow = []
    for _ in range(steps):
        slow = md.step(xd, vd, slow)
    d = np.abs(xd.cpu().numpy() - xo)
    d = np.minimum(d, np.abs(d - box[None, :]))
    assert d.max() < 1e-9
    assert np.abs(vd.cpu().numpy() - vo).max() < 1e-9
    en = md.bonded_energies()
    assert en[2] + en[3] == pytest.approx(e_pot, rel=1e-9)
    if fused:       # per-kind force arrays of the last positions, on request
        assert np.abs((fout[0] + fout[1]).cpu().numpy() - fo_).max() < 1e-7 * np.abs(fo_).max()


@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_fused_inner_step_equals_separate_launches(real):
    """hymd_bonded_inner_step (forces + kicks + drift in one pass) against the four separate launches
    per inner step on the same inputs, all three kinds present, odd respa_inner (result lands in the
    alternate position buffer and is copied back)."""
    from hymd_b200.force import BondedTopology
    from hymd_b200.md import RespaMD
    rng = np.random.default_rng(12)
    box = np.array([4.0, 5.0, 4.5])
    r, a2, a3, a4 = chains(rng, 150, 9, box, real)
    n = len(r)
    coeff = np.zeros((len(a4), 6, 5))
    coeff[:, 0] = rng.normal(size=(len(a4), 5))
    coeff[:, 1] = rng.uniform(-np.pi, np.pi, size=(len(a4), 5))
    topo = BondedTopology(n, bonds=(a2, a2 + 1, np.full(len(a2), 0.47), np.full(len(a2), 1250.0)),
                          angles=(a3, a3 + 1, a3 + 2, np.full(len(a3), 2.0), np.full(len(a3), 25.0)),
                          dihedrals=(a4, a4 + 1, a4 + 2, a4 + 3, coeff, np.zeros(len(a4), dtype=int)),
                          device=DEVICE if DEVICE != "cuda" else None)
    v = rng.normal(scale=0.15, size=(n, 3)).astype(real)
    slow = dev(rng.normal(scale=50.0, size=(n, 3)), real)
    out = {}
    for fused in (True, False):
        md = RespaMD(lambda x: [slow], box, 72.0, 0.004, respa_inner=3, topology=topo, fused=fused)
        xd, vd = dev(r, real), dev(v, real)
        f = [slow]
        for _ in range(4):
            f = md.step(xd, vd, f)
        out[fused] = (xd.cpu().numpy().astype(np.float64), vd.cpu().numpy().astype(np.float64), md.bonded_energies())
    eps = np.finfo(real).eps
    d = np.abs(out[True][0] - out[False][0])
    d = np.minimum(d, np.abs(d - box[None, :]))
    assert d.max() <= 64 * eps * box.max()
    assert np.abs(out[True][1] - out[False][1]).max() <= 64 * eps * np.abs(out[False][1]).max()
    for k in (2, 3, 4):
        assert out[True][2][k] == pytest.approx(out[False][2][k], rel=1e-6 if real == np.float32 else 1e-11)


@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_respa_md_with_cbt_dihedrals_fused_equals_separate_launches(real):
    """A topology with dih_type 1 through RespaMD: the fused inner step runs the per-particle kernel variant that
    carries the bending term (inner_step_kernel<real, true>, whatever ``cta`` says) and must follow the separate
    launches (dihedral kernel + bending pass + kick / drift) to rounding; the dihedral energy it reports is the
    oracle's for the final positions."""
    from hymd_b200.force import BondedTopology
    from hymd_b200.md import RespaMD
    rng = np.random.default_rng(44)
    box = np.array([4.0, 5.0, 4.5])
    n = 60
    r = np.mod(np.cumsum(rng.normal(scale=0.25, size=(n, 3)), axis=0) + 2.0, box).astype(real)
    a2, a3, a4 = np.arange(n - 1), np.arange(n - 2), np.arange(n - 3)
    coeff = np.zeros((len(a4), 6, 5))
    coeff[:, 0] = rng.uniform(0.5, 2.0, size=(len(a4), 5))
    coeff[:, 4] = rng.uniform(20, 40, size=(len(a4), 5))
    dt4 = np.ones(len(a4), dtype=int)
    dt4[::5] = 0                      # a few plain cosine-series dihedrals among them
    last = np.zeros(len(a4), dtype=int)
    last[-1] = 1
    topo = BondedTopology(n, bonds=(a2, a2 + 1, np.full(n - 1, 0.47), np.full(n - 1, 1250.0)),
                          angles=(a3, a3 + 1, a3 + 2, np.full(n - 2, 2.0), np.full(n - 2, 25.0)),
                          dihedrals=(a4, a4 + 1, a4 + 2, a4 + 3, coeff, dt4, last),
                          device=DEVICE if DEVICE != "cuda" else None)
    assert topo.n_cbt == int((dt4 == 1).sum())
    v = rng.normal(scale=0.1, size=(n, 3)).astype(real)
    out = {}
    for fused in (True, False):
        md = RespaMD(lambda x: [torch.zeros_like(x)], box, 72.0, 0.002, respa_inner=3, topology=topo, fused=fused)
        xd, vd = dev(r, real), dev(v, real)
        f = [torch.zeros_like(xd)]
        for _ in range(3):
            f = md.step(xd, vd, f)
        out[fused] = (xd.cpu().numpy().astype(np.float64), vd.cpu().numpy().astype(np.float64), md.bonded_energies())
    eps = np.finfo(real).eps
    d = np.abs(out[True][0] - out[False][0])
    d = np.minimum(d, np.abs(d - box[None, :]))
    assert d.max() <= 256 * eps * box.max()
    assert np.abs(out[True][1] - out[False][1]).max() <= 256 * eps * np.abs(out[False][1]).max()
    for k in (2, 3, 4):
        assert out[True][2][k] == pytest.approx(out[False][2][k], rel=2e-5 if real == np.float32 else 1e-10)
    # energies belong to the positions of the last inner force evaluation = the final positions
    fo_, eo = bo.compute_dihedral_forces(out[True][0], box, a4, a4 + 1, a4 + 2, a4 + 3, coeff, dt4, last)
    assert out[True][2][4] == pytest.approx(eo, rel=1e-5 if real == np.float32 else 1e-10)
    # the standalone entry point with the per-kind force arrays: bending + torsion land in the dihedral array
    x = dev(out[True][0], real)
    f3 = [torch.zeros_like(x) for _ in range(3)]
    vv = dev(v, real)
    topo.inner_step(x, None, vv, box, 72.0, 0.002, 1, 0.0, force_out=f3, cta=1)
    assert np.abs(f3[2].cpu().numpy() - fo_).max() < (1e-5 if real == np.float32 else 1e-10) * np.abs(fo_).max()


def test_fused_step_edge_cases():
    """Empty topology (monatomic system), a single particle, and a 129-particle chain whose terms straddle
    the CTA boundary, through the fused inner step in per-particle and CTA-cooperative mode."""
    from hymd_b200.force import BondedTopology
    box = np.array([50.0, 50.0, 50.0])
    rng = np.random.default_rng(3)
    for n in (1, 7):
        topo = BondedTopology(n, device=DEVICE if DEVICE != "cuda" else None)
        x = dev(rng.random((n, 3)) * 5, np.float64)
        v = dev(rng.normal(size=(n, 3)), np.float64)
        for cta in (0, 1):
            x1, v1, x2 = x.clone(), v.clone(), torch.empty_like(x)
            res = topo.inner_step(x1, x2, v1, box, 72.0, 0.01, 2, 0.01, cta=cta)
            assert float(res.abs().max()) == 0.0 and torch.equal(v1, v)          # no forces: pure drift
            assert torch.allclose(x2, x + 0.01 * v, rtol=0, atol=1e-14)
    n = 129
    steps = rng.normal(size=(n - 1, 3))
    steps *= 0.3 / np.linalg.norm(steps, axis=1)[:, None]
    r = 10.0 + np.concatenate([np.zeros((1, 3)), np.cumsum(steps, 0)])
    a2, a3, a4 = np.arange(n - 1), np.arange(n - 2), np.arange(n - 3)
    coeff = np.zeros((n - 3, 6, 5))
    coeff[:, 0, 1:3] = [1.0, -0.5]
    terms = dict(bonds=(a2, a2 + 1, np.full(n - 1, 0.25), np.full(n - 1, 500.0)),
                 angles=(a3, a3 + 1, a3 + 2, np.full(n - 2, 2.0), np.full(n - 2, 20.0)),
                 dihedrals=(a4, a4 + 1, a4 + 2, a4 + 3, coeff, np.zeros(n - 3, dtype=int)))
    topo = BondedTopology(n, device=DEVICE if DEVICE != "cuda" else None, **terms)
    fb, eb, _ = bo.compute_bond_forces(r, box, *terms["bonds"])
    fa, ea, _ = bo.compute_angle_forces(r, box, *terms["angles"])
    fd, ed = bo.compute_dihedral_forces(r, box, *terms["dihedrals"])
    x = dev(r, np.float64)
    for cta in (0, 1, 2):
        f = [torch.zeros((n, 3), dtype=torch.float64, device=DEVICE) for _ in range(3)]
        res = topo.inner_step(x, None, torch.zeros_like(x), box, 72.0, 0.0, 0, 0.0, force_out=f, cta=cta).clone()
        for got, want in zip(f, (fb, fa, fd)):
            assert np.abs(got.cpu().numpy() - want).max() <= 1e-10 * np.abs(want).max()
        np.testing.assert_allclose(res[:, 0].cpu().numpy(), [eb, ea, ed], rtol=1e-11)
    topo.set_cta(0)


@pytest.mark.parametrize("real", [np.float32, np.float64])
@pytest.mark.parametrize("tile", [128, 512])
def test_cta_cooperative_evaluation_matches_the_per_particle_result(real, tile, monkeypatch):
    """hymd_bonded_set_cta(1): each CTA evaluates every term touching its 128 particles once into shared
    memory and the particles gather their slots.  Same additions in the same order => the forces equal the
    per-particle kernels' up to FMA contraction across the inlined term evaluation (bit-identical when
    the same source runs on the CPU, tests/test_md_host_emulation.py); energies agree to rounding.  Branched molecules (DPPC
    topology of the reference's fixture), chains straddling CTA boundaries, all three kinds, a solvent
    tail without terms; then a fused rRESPA run with the switch on.  ``tile`` = particles per CTA
    (HYMD_B200_BONDED_TILE: 512 = four particles per thread)."""
    monkeypatch.setenv("HYMD_B200_BONDED_TILE", str(tile))
    from hymd_b200.force import BondedTopology
    from hymd_b200.md import RespaMD
    rng = np.random.default_rng(77)
    box = np.array([6.0, 5.0, 7.0])
    # 90 copies of the 12-bead DPPC graph (test/conftest.py dppc_single) + 140 chains of 9 + 300 solvent
    dppc_b = np.array([[0, 1], [1, 2], [2, 3], [2, 4], [3, 8], [4, 5], [5, 6], [6, 7], [8, 9], [9, 10], [10, 11]])
    dppc_a = np.array([[1, 2, 3], [1, 2, 4], [3, 8, 9], [2, 4, 5], [4, 5, 6], [5, 6, 7], [8, 9, 10], [9, 10, 11]])
    n_d = 90
    r_d = (G["dppc/r"][None] - G["dppc/r"].mean(axis=0) + (rng.random((n_d, 1, 3)) * box)).reshape(-1, 3)
    b2 = (dppc_b[None] + 12 * np.arange(n_d)[:, None, None]).reshape(-1, 2)
    b3 = (dppc_a[None] + 12 * np.arange(n_d)[:, None, None]).reshape(-1, 3)
    r_c, a2, a3, a4 = chains(rng, 140, 9, box, np.float64)
    off = 12 * n_d
    b2 = np.concatenate([b2, np.stack([a2, a2 + 1], 1) + off])
    b3 = np.concatenate([b3, np.stack([a3, a3 + 1, a3 + 2], 1) + off])
    b4 = np.stack([a4, a4 + 1, a4 + 2, a4 + 3], 1) + off
    r = np.mod(np.concatenate([r_d, r_c, rng.random((300, 3)) * box]), box).astype(real)
    n = len(r)
    r0, k2 = 0.47 + 0.05 * rng.random(len(b2)), 1250.0 * (0.5 + rng.random(len(b2)))
    t0, k3 = np.radians(rng.choice([120.0, 180.0], size=len(b3))), 25.0 * (0.5 + rng.random(len(b3)))
    coeff = np.zeros((len(b4), 6, 5))
    coeff[:, 0] = rng.normal(size=(len(b4), 5))
    coeff[:, 1] = rng.uniform(-np.pi, np.pi, size=(len(b4), 5))
    dt4 = np.zeros(len(b4), dtype=int)
    dt4[::4] = 2
    coeff[::4, 0, 0], coeff[::4, 0, 1] = -0.5, 30.0
    topo = BondedTopology(n, bonds=(b2[:, 0], b2[:, 1], r0, k2), angles=(b3[:, 0], b3[:, 1], b3[:, 2], t0, k3),
                          dihedrals=(b4[:, 0], b4[:, 1], b4[:, 2], b4[:, 3], coeff, dt4),
                          device=DEVICE if DEVICE != "cuda" else None)
    pos = dev(r, real)
    ref = {}
    modes = (0, 1, 2, 3) if (DEVICE == "cpu" or os.environ.get("HYMD_TEST_CTA3") == "1") else (0, 1, 2)
    for cta in modes:      # mode 3 has not run on a GPU yet: tools/gpu_md_next.sh sets HYMD_TEST_CTA3=1
        topo.set_cta(cta)
        for kind in (2, 3, 4):
            f = torch.full((n, 3), 3.0, dtype=pos.dtype, device=DEVICE)
            res = topo.forces(kind, pos, box, f).clone()
            if cta == 0:
                ref[kind] = (f.clone(), res)
            else:
                fs = float(ref[kind][0].abs().max())
                assert float((f - ref[kind][0]).abs().max()) <= 4 * np.finfo(real).eps * fs, f"kind {kind}"
                if DEVICE == "cpu":
                    assert torch.equal(f, ref[kind][0])
                scale = float(ref[kind][1].abs().max()) or 1.0
                assert float((res - ref[kind][1]).abs().max()) <= 1e-12 * scale
    fo_, eo, pro = bo.compute_angle_forces(r, box, b3[:, 0], b3[:, 1], b3[:, 2], t0, k3)
    assert np.abs(ref[3][0].cpu().numpy() - fo_).max() <= FTOL[real] * np.abs(fo_).max()
    # fused rRESPA steps with the switch on == off
    v = rng.normal(scale=0.15, size=(n, 3)).astype(real)
    out = {}
    for cta in modes:
        topo.set_cta(cta)
        md = RespaMD(lambda x: [], box, 72.0, 0.004, respa_inner=4, topology=topo, cta=cta)
        xd, vd = dev(r, real), dev(v, real)
        for _ in range(3):
            md.step(xd, vd, [])
        out[cta] = (xd.clone(), vd.clone(), md.bonded_energies())
    eps = np.finfo(real).eps
    for cta in modes[1:]:
        d = (out[0][0] - out[cta][0]).abs().cpu().numpy()
        d = np.minimum(d, np.abs(d - box[None, :].astype(real)))
        assert d.max() <= 256 * eps * box.max()
        assert float((out[0][1] - out[cta][1]).abs().max()) <= 256 * eps * float(out[0][1].abs().max())
        for k in (2, 3, 4):
            assert out[cta][2][k] == pytest.approx(out[0][2][k], rel=1e-5 if real == np.float32 else 1e-10)
    topo.set_cta(False)


def test_barostat_rescale_then_fields_match_the_oracle_in_the_new_box():
    """NPT step (main.py:889-935): hymd_b200.barostat.isotropic computes the pressure on the device,
    rescales box and positions in place and tells the context the new box (hymd_ctx_set_box) instead of
    re-running initialize_pm; the next field-force cycle must equal the oracle's in the rescaled box."""
    import types as pytypes

    from conftest import make_config
    from gpu_common import GpuRun, OracleRun, rel_err
    from hymd_b200 import barostat as B
    from hymd_b200 import field as F
    from oracle import field_oracle as fo
    rng = np.random.default_rng(21)
    n, mesh, box = 4000, [16, 12, 10], np.array([4.0, 5.0, 6.0], dtype=np.float32)
    names = [("A", "B")[i % 2] for i in range(n)]
    cfg = make_config(names, n, mesh, box, chi=[("A", "B", 15.0)], dtype=np.float64)
    cfg.n_b, cfg.tau_p, cfg.target_pressure = 1, 0.05, pytypes.SimpleNamespace(P_L=1.0, P_N=1.0)
    cfg.barostat = "isotropic"          # Hamiltonian._setup then keeps rho0 / a fixed (hamiltonian.py:41-47)
    cfg.rho0 = cfg.a = n / float(np.prod(box.astype(np.float64)))
    types_ = np.array([cfg.name_to_type_map[t] for t in names], dtype=np.int32)
    pos = (rng.random((n, 3)) * box).astype(np.float64)
    vel = rng.normal(scale=0.2, size=(n, 3))
    g = GpuRun(cfg, pos, types_, compute_potential=True)
    o = OracleRun(cfg, pos, types_)
    bond_pr, angle_pr = np.array([1.0, -2.0, 0.5]), np.array([0.25, 0.5, -1.0])
    p_ref = fo.comp_pressure(o.st, o.h, vel, o.cfg, bond_pr, angle_pr)
    P = np.average(p_ref[-3:-1]) * 16.61
    alpha = (1.0 - cfg.time_step * cfg.respa_inner * cfg.n_b / cfg.tau_p * 4.6e-5 * (1.0 - P)) ** (1 / 3)
    assert abs(alpha - 1.0) > 1e-4            # a rescale the parity tolerance can see
    stuff = (g.pm, None, None, None)
    box_before = np.array(cfg.box_size, dtype=np.float64)
    res, change = B.isotropic(None, stuff, g.phi, g.phi_q, g.psi, g.h, g.pos, dev(vel, np.float64), cfg,
                              g.phi_fourier, g.phi_laplacian, g.phi_transfer, bond_pr, angle_pr, 3, None)
    assert change and res is stuff
    np.testing.assert_allclose(np.asarray(cfg.box_size, dtype=np.float64), box_before * alpha, rtol=1e-6)
    np.testing.assert_allclose(g.pos.cpu().numpy(), pos * alpha, rtol=1e-9)
    layouts = [g.pm.decompose(None) for _ in range(cfg.n_types)]
    F.update_field(g.phi, g.phi_laplacian, g.phi_transfer, layouts, g.force_mesh, g.h, g.pm, g.pos, g.types,
                   cfg, g.v_ext, g.phi_fourier, g.v_ext_fourier, cfg.m)
    F.compute_field_force(layouts, g.pos, g.force_mesh, g.force, g.types, cfg.n_types)
    o2 = OracleRun(cfg, g.pos.cpu().numpy(), types_)
    assert rel_err(g.forces(), o2.force) < 1e-10


def test_f32_math_inner_step_stays_within_the_fp32_tolerance():
    """hymd_bonded_set_math(1): float arithmetic for bonds and angles in the per-particle fused step of the
    fp32 build, against the default double arithmetic: forces within 1e-5 of the largest force (north_star
    tolerance of the fp32 build; measured 3e-7 on the CPU), same trajectory to that accuracy."""
    from hymd_b200.force import BondedTopology
    rng = np.random.default_rng(41)
    box = np.array([5.0, 6.0, 7.0])
    r, a2, a3, a4 = chains(rng, 200, 10, box, np.float32)
    n = len(r)
    coeff = np.zeros((len(a4), 6, 5))
    coeff[:, 0] = rng.normal(size=(len(a4), 5))
    topo = BondedTopology(n, bonds=(a2, a2 + 1, 0.47 + 0.05 * rng.random(len(a2)), np.full(len(a2), 1250.0)),
                          angles=(a3, a3 + 1, a3 + 2, np.radians(rng.choice([120.0, 180.0], size=len(a3))),
                                  np.full(len(a3), 25.0)),
                          dihedrals=(a4, a4 + 1, a4 + 2, a4 + 3, coeff, np.zeros(len(a4), dtype=int)),
                          device=DEVICE if DEVICE != "cuda" else None)
    topo.set_cta(0)
    x = dev(r, np.float32)
    v0 = rng.normal(scale=0.15, size=(n, 3)).astype(np.float32)
    res = {}
    for f32 in (False, True):
        topo.set_math(f32)
        f = [torch.zeros((n, 3), dtype=torch.float32, device=DEVICE) for _ in range(3)]
        v, x2 = dev(v0, np.float32), torch.empty_like(x)
        e = topo.inner_step(x, x2, v, box, 72.0, 0.01, 2, 0.01, force_out=f).clone()
        res[f32] = ([t.cpu().numpy().astype(np.float64) for t in f], v.cpu().numpy(), x2.cpu().numpy(), e.cpu().numpy())
    # the standalone bond / angle kernels follow the same switch
    for kind in (2, 3):
        out = {}
        for f32 in (False, True):
            topo.set_math(f32)
            fk = torch.zeros((n, 3), dtype=torch.float32, device=DEVICE)
            ek = topo.forces(kind, x, box, fk).clone()
            out[f32] = (fk.cpu().numpy().astype(np.float64), ek.cpu().numpy())
        assert np.abs(out[True][0] - out[False][0]).max() <= 1e-5 * np.abs(out[False][0]).max()
        assert out[True][1][0] == pytest.approx(out[False][1][0], rel=1e-5)
        assert np.abs(out[True][0] - res[True][0][kind - 2]).max() <= 1e-6 * np.abs(out[False][0]).max()
    topo.set_math(False)
    scale = max(np.abs(t).max() for t in res[False][0])
    for k in range(3):
        assert np.abs(res[True][0][k] - res[False][0][k]).max() <= 1e-5 * scale
    assert np.array_equal(res[True][0][2], res[False][0][2])          # dihedrals keep the double evaluator
    assert np.abs(res[True][1] - res[False][1]).max() <= 1e-5 * np.abs(res[False][1]).max()
    np.testing.assert_allclose(res[True][3][:2, 0], res[False][3][:2, 0], rtol=1e-5)

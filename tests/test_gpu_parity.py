"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): per-particle field forces and field energies agree
within 1e-5 relative for the fp32 build and 1e-10 for the fp64 build; relative error is
max|a-b| / max|b|.  fp32 results are compared against the fp64 oracle evaluated on the same
(float32-valued) inputs."""
import numpy as np
import pytest
import torch

from conftest import make_config

pytestmark = pytest.mark.gpu

TOL = {np.float32: 1e-5, np.float64: 1e-10}
CHI3 = [("A", "B", 9.6754032616815161), ("A", "C", -13.2596290315913623),
        ("B", "C", 0.3852001771213374)]


def _system(n, mesh, box, dtype, seed=0, names=("A", "B", "C"), chi=CHI3, coulomb=False,
            hamiltonian="DefaultWithChi", m=None):
    rng = np.random.default_rng(seed)
    box = np.asarray(box, dtype=np.float32)
    pos = (rng.uniform(0, 1, size=(n, 3)) * box).astype(dtype)
    pos = np.minimum(pos, np.nextafter(box.astype(dtype), 0).astype(dtype))
    tnames = [names[i % len(names)] for i in range(n)]
    cfg = make_config(tnames, n, mesh, box, chi=chi, dtype=dtype, hamiltonian=hamiltonian,
                      coulombtype="PIC_Spectral" if coulomb else None,
                      dielectric_const=80.0 if coulomb else None, m=m)
    types = np.array([cfg.name_to_type_map[t] for t in tnames], dtype=np.int32)
    q = None
    if coulomb:
        q = rng.choice([-1.0, 0.0, 1.0], size=n)
        q -= q.mean()
        q = q.astype(dtype)
    return cfg, pos, types, q


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mesh", [[24, 24, 24], [16, 12, 10], [9, 12, 10], [40, 24, 72], [5, 5, 5]])
def test_paint_matches_oracle(dtype, mesh):
    """Raw CIC densities / dV (field.py:574-575) for every type, and mass conservation
    (test_hamiltonian.py:105-109)."""
    from gpu_common import GpuRun, rel_err
    from hymd_b200 import _lib
    from oracle import pm_oracle as pmo
    from oracle.field_oracle import volume_per_cell
    cfg, pos, types, _ = _system(3000, mesh, [4.0, 5.0, 6.0], dtype)
    g = GpuRun(cfg, pos, types)
    dv = volume_per_cell(cfg)
    for t in range(cfg.n_types):
        got = g.pm._view(_lib.FIELD_PHI, t, 0, "real").cpu().numpy()
        want = pmo.cic_paint(pos[types == t].astype(np.float64), 1.0, mesh, cfg.box_size) / dv
        assert rel_err(got, want) < (1e-6 if dtype == np.float32 else 1e-13)
        assert got.sum(dtype=np.float64) * dv == pytest.approx((types == t).sum(), rel=1e-6 if dtype == np.float32 else 1e-13)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mesh,n", [([24, 24, 24], 10000), ([16, 12, 10], 700), ([9, 12, 10], 500),
                                    ([32, 40, 64], 20000), ([64, 48, 40], 30000),
                                    ([128, 16, 24], 8000), ([256, 10, 12], 6000),
                                    # power-of-two square (y,z) planes: one-pass plane transforms
                                    ([16, 16, 16], 3000), ([32, 64, 64], 20000), ([64, 32, 32], 9000),
                                    ([128, 128, 128], 60000), ([16, 256, 256], 30000)])
def test_field_forces_match_oracle(dtype, mesh, n):
    from gpu_common import GpuRun, OracleRun, rel_err
    cfg, pos, types, _ = _system(n, mesh, [4.0, 5.0, 6.0], dtype, seed=1)
    g = GpuRun(cfg, pos, types)
    o = OracleRun(cfg, pos, types)
    assert rel_err(g.forces(), o.force) < TOL[dtype]
    for t in range(cfg.n_types):
        for d in range(3):
            assert rel_err(g.force_mesh[t][d].value.cpu().numpy(), o.st.force_mesh[t][d]) < TOL[dtype]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind,mesh", [("DefaultNoChi", [20, 24, 28]), ("SquaredPhi", [20, 24, 28]),
                                       ("DefaultWithChi", [20, 24, 28]),
                                       ("DefaultWithChi", [32, 24, 28]), ("DefaultNoChi", [64, 20, 18]),
                                       ("DefaultWithChi", [32, 64, 64]), ("DefaultWithChi", [16, 128, 128])])
def test_energies_and_potentials_match_oracle(dtype, kind, mesh):
    """compute_potential=True path: filtered densities, v_ext and the field energy
    (field.py:578, 615-616, 692-693); power-of-two Nx runs the fused x-line kernel."""
    from gpu_common import GpuRun, OracleRun, rel_err
    cfg, pos, types, _ = _system(5000, mesh, [4.0, 5.0, 6.0], dtype, seed=2,
                                 hamiltonian=kind)
    rng = np.random.default_rng(9)
    vel = rng.normal(size=pos.shape).astype(dtype)
    g = GpuRun(cfg, pos, types, compute_potential=True)
    o = OracleRun(cfg, pos, types)
    tol = TOL[dtype]
    for t in range(cfg.n_types):
        assert rel_err(g.phi[t].value.cpu().numpy(), o.st.phi[t]) < tol
        assert rel_err(g.phi_fourier[t].value.cpu().numpy(), o.st.phi_fourier[t]) < tol
        # v_ext ~ fluctuation around zero of terms of size A*rho0: compare on that scale
        scale = np.abs(o.st.v_ext[t]).max() + 1.0 / cfg.kappa
        assert np.abs(g.v_ext[t].value.cpu().numpy() - o.st.v_ext[t]).max() / scale < tol
    e_g = g.energies(torch.as_tensor(vel, device="cuda"))
    e_o = o.energies(vel)
    # the field energy is a sum of squared fluctuations; compare on the scale N/(2 kappa)
    escale = max(abs(e_o[0]), 0.5 * len(pos) / cfg.kappa * 1e-2)
    assert abs(e_g[0] - e_o[0]) / escale < tol
    assert e_g[1] == pytest.approx(e_o[1], rel=1e-6 if dtype == np.float32 else 1e-12)
    assert e_g[2] == 0.0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mesh", [[24, 20, 28], [32, 20, 28], [128, 12, 16], [32, 32, 32], [64, 128, 128]])
def test_pme_matches_oracle(dtype, mesh):
    from gpu_common import GpuRun, OracleRun, rel_err
    from hymd_b200.field import compute_self_energy_q
    from oracle.field_oracle import compute_self_energy_q as self_o
    cfg, pos, types, q = _system(6000, mesh, [4.0, 5.0, 6.0], dtype, seed=3, coulomb=True)
    cfg.self_energy = self_o(cfg, q)
    assert compute_self_energy_q(cfg, torch.as_tensor(q, device="cuda")) == pytest.approx(cfg.self_energy, rel=1e-6)
    g = GpuRun(cfg, pos, types, charges=q)
    o = OracleRun(cfg, pos, types, charges=q)
    tol = TOL[dtype]
    assert rel_err(g.eforces(), o.elec_forces) < tol
    assert rel_err(g.forces(), o.force) < tol
    for d in range(3):
        assert rel_err(g.elec_field[d].value.cpu().numpy(), o.st.elec_field[d]) < tol
    vel = np.zeros_like(pos)
    e_g = g.energies(vel)
    e_o = o.energies(vel)
    assert rel_err(g.psi.value.cpu().numpy(), o.st.psi) < tol
    assert rel_err(g.phi_q.value.cpu().numpy(), o.st.phi_q) < (1e-6 if dtype == np.float32 else 1e-12)
    assert rel_err(g.phi_q_fourier.value.cpu().numpy(), o.st.phi_q_fourier) < tol
    assert abs(e_g[2] - e_o[2]) / max(abs(e_o[2]), abs(cfg.self_energy)) < tol


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bitwise_deterministic_and_order_independent(dtype):
    """Same particles in a different caller order (and a second run): densities are bitwise
    identical, per-particle forces are bitwise identical after un-permuting."""
    from gpu_common import GpuRun
    from hymd_b200 import _lib
    cfg, pos, types, _ = _system(20000, [24, 24, 24], [3.0, 3.0, 3.0], dtype, seed=4)
    # make some cells crowded
    pos[:2000] = pos[0] + (np.random.default_rng(1).uniform(0, 0.05, size=(2000, 3))).astype(dtype)
    pos = np.mod(pos, cfg.box_size.astype(dtype)).astype(dtype)
    a = GpuRun(cfg, pos, types)
    phi_a = [a.pm._view(_lib.FIELD_PHI, t, 0, "real").clone() for t in range(3)]
    fa = a.forces()
    b = GpuRun(cfg, pos, types)
    assert np.array_equal(fa, b.forces())
    perm = np.random.default_rng(7).permutation(len(pos))
    c = GpuRun(cfg, pos[perm], types[perm])
    for t in range(3):
        assert torch.equal(phi_a[t], c.pm._view(_lib.FIELD_PHI, t, 0, "real"))
    assert np.array_equal(fa[perm], c.forces())


def test_numpy_fortran_order_inputs_and_inplace_outputs():
    """main.py hands over host numpy arrays, Fortran-ordered when molecules exist
    (main.py:500-506); outputs are written in place into the caller's array."""
    from gpu_common import GpuRun
    cfg, pos, types, q = _system(4000, [16, 16, 16], [3.0, 3.0, 3.0], np.float64, seed=5, coulomb=True)
    ref = GpuRun(cfg, pos, types, charges=q)
    posf = np.asfortranarray(pos)
    g = GpuRun(cfg, posf, types.astype(np.int64), charges=q, as_numpy=True)
    assert isinstance(g.force, np.ndarray)
    assert np.array_equal(g.force, ref.forces())
    assert np.array_equal(g.elec_forces, ref.eforces())


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_edge_cases(dtype):
    """Empty type, a single particle, particles on the box faces and outside [0, L)."""
    from gpu_common import GpuRun, OracleRun, rel_err
    box = np.array([3.0, 3.0, 3.0], dtype=np.float32)
    mesh = [12, 12, 12]
    # type "B" exists in the config but has no particles
    rng = np.random.default_rng(6)
    n = 300
    pos = (rng.uniform(0, 1, size=(n, 3)) * box).astype(dtype)
    pos[0] = [0.0, 0.0, 0.0]
    pos[1] = [3.0, 3.0, 3.0]                 # == L: wraps onto vertex 0
    pos[2] = [-0.25, 3.5, 7.0]               # outside the box: periodic wrap (CIC wraps mod N)
    pos[3] = np.nextafter(np.asarray([3.0, 3.0, 3.0], dtype=dtype), 0)
    names = ["A"] * n
    cfg = make_config(names + ["B"], n, mesh, box, chi=[("A", "B", 5.0)], dtype=dtype)
    types = np.zeros(n, dtype=np.int32)
    g = GpuRun(cfg, pos, types)
    o = OracleRun(cfg, pos, types)
    assert rel_err(g.forces(), o.force) < TOL[dtype]
    # single particle: forces ~ 0 by symmetry of its own filtered density, must be finite
    cfg1 = make_config(["A"], 1, mesh, box, dtype=dtype, hamiltonian="DefaultNoChi")
    g1 = GpuRun(cfg1, pos[5:6], np.zeros(1, dtype=np.int32))
    o1 = OracleRun(cfg1, pos[5:6], np.zeros(1, dtype=np.int32))
    assert np.isfinite(g1.forces()).all()
    scale = np.abs(np.stack([np.stack(fm) for fm in o1.st.force_mesh])).max()
    assert np.abs(g1.forces() - o1.force).max() / scale < TOL[dtype]


def test_per_type_paint_mass_and_shared_potential_rows():
    """config.m per-type paint masses (field.py:574) and types whose interaction rows coincide
    (DefaultNoChi: all of them) share one force-mesh triple."""
    from gpu_common import GpuRun, OracleRun, rel_err
    cfg, pos, types, _ = _system(3000, [16, 16, 16], [3.0, 3.0, 3.0], np.float64, seed=8,
                                 hamiltonian="DefaultNoChi", chi=(), m=[1.0, 2.5, 0.5])
    g = GpuRun(cfg, pos, types)
    o = OracleRun(cfg, pos, types)
    assert g.pm.status()["potential_rows"] == 1
    assert rel_err(g.forces(), o.force) < 1e-10


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_consecutive_steps_reuse_previous_order(dtype):
    """MD-like sequence on ONE context: positions drift between calls, the types tensor is the
    same object, so the binning starts from the previous cell order (HYMD_SORT_REUSE_ORDER).
    Results must be bitwise identical to a cold context and match the oracle."""
    import torch
    from gpu_common import GpuRun, OracleRun, rel_err
    from hymd_b200 import field as F
    cfg, pos, types, q = _system(15000, [32, 24, 20], [4.0, 5.0, 6.0], dtype, seed=12, coulomb=True)
    rng = np.random.default_rng(3)
    g = GpuRun(cfg, pos, types, charges=q)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    box = cfg.box_size.astype(dtype)
    layouts = [g.pm.decompose(None) for _ in range(cfg.n_types)]
    for step in range(3):
        pos = np.mod(pos + rng.normal(scale=0.05, size=pos.shape).astype(dtype), box).astype(dtype)
        pos[pos >= box] = 0
        pos_d = torch.as_tensor(pos, dtype=tdt, device="cuda")
        F.update_field(g.phi, g.phi_laplacian, g.phi_transfer, layouts, g.force_mesh, g.h, g.pm,
                       pos_d, g.types, cfg, g.v_ext, g.phi_fourier, g.v_ext_fourier, cfg.m)
        F.compute_field_force(layouts, pos_d, g.force_mesh, g.force, g.types, cfg.n_types)
        F.update_field_force_q(g.q, g.phi_q, g.phi_q_fourier, g.psi, None, None, g.elec_field,
                               g.elec_forces, g.pm.decompose(None), g.h, g.pm, pos_d, cfg)
    torch.cuda.synchronize()
    cold = GpuRun(cfg, pos, types, charges=q)
    assert np.array_equal(g.forces(), cold.forces())
    assert np.array_equal(g.eforces(), cold.eforces())
    o = OracleRun(cfg, pos, types, charges=q)
    assert rel_err(g.forces(), o.force) < TOL[dtype]
    assert rel_err(g.eforces(), o.elec_forces) < TOL[dtype]


def test_domain_decomposition_returns_cell_order_single_gpu():
    """domain_decomposition (field.py:1115-1178) permutes every per-particle array identically;
    this implementation hands them back in mesh-cell order with molecules contiguous and in
    their internal order.  Forces computed on the permuted arrays are the same particles'
    forces; numpy in -> numpy out, torch in -> torch out."""
    from gpu_common import GpuRun
    from hymd_b200 import field as F
    cfg, pos, types, _ = _system(6000, [16, 16, 16], [3.0, 3.0, 3.0], np.float64, seed=21)
    n = len(pos)
    rng = np.random.default_rng(5)
    # molecules of 1..6 consecutive atoms, atoms of a molecule close to its first atom
    sizes = rng.integers(1, 7, size=n)
    mol = np.repeat(np.arange(n), sizes)[:n].astype(np.int32)
    first = np.concatenate([[0], np.nonzero(np.diff(mol))[0] + 1])
    start_of = first[np.searchsorted(first, np.arange(n), side="right") - 1]
    pos = np.mod(pos[start_of] + rng.normal(scale=0.05, size=pos.shape), cfg.box_size.astype(np.float64))
    gid = np.arange(n, dtype=np.int64)
    bonds = rng.integers(-1, 5, size=(n, 3)).astype(np.int32)
    ref = GpuRun(cfg, pos, types)
    out = F.domain_decomposition(pos, ref.pm, types, gid, molecules=mol, bonds=bonds)
    p2, t2, g2, b2, m2 = out
    assert all(isinstance(a, np.ndarray) for a in out)
    assert np.array_equal(np.sort(g2), gid)
    assert np.array_equal(p2, pos[g2]) and np.array_equal(t2, types[g2])
    assert np.array_equal(b2, bonds[g2]) and np.array_equal(m2, mol[g2])
    # molecules contiguous, atoms in their original order
    change = np.nonzero(np.diff(m2))[0] + 1
    assert len(change) + 1 == len(np.unique(mol))
    assert (np.diff(g2)[np.diff(m2) == 0] == 1).all()
    # first atoms are in mesh-cell order
    firsts = np.concatenate([[0], change])
    c = np.floor(p2[firsts] * 16 / 3.0).astype(np.int64) % 16
    key = (c[:, 0] * 16 + c[:, 1]) * 16 + c[:, 2]
    assert (np.diff(key) >= 0).all()
    g = GpuRun(cfg, p2, t2)
    assert np.array_equal(g.forces(), ref.forces()[g2])
    # torch tensors in -> torch tensors out, same permutation
    outt = F.domain_decomposition(torch.as_tensor(pos, device="cuda"), ref.pm,
                                  torch.as_tensor(gid, device="cuda"))
    assert isinstance(outt[0], torch.Tensor) and outt[0].is_cuda
    assert np.array_equal(np.sort(outt[1].cpu().numpy()), gid)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("env", [{"HYMD_B200_GRAD2": "0"}, {"HYMD_B200_PLANE_TILES": "2"},
                                 {"HYMD_B200_PLANE_TILES": "3"},
                                 {"HYMD_B200_PLANE_TILES": "2", "HYMD_B200_ROW_TMA": "3", "HYMD_B200_GRAD2": "0"},
                                 {"HYMD_B200_NO_FUSED": "1"}, {"HYMD_B200_NO_PLANE": "1"}])
def test_kernel_variants_match_oracle(dtype, env, monkeypatch):
    """The tuning switches select other kernels for the same arithmetic: three force spectra per
    potential row instead of two (+ k_y, k_z applied by the plane c2r), double-buffered column
    tiles with row inputs / outputs staged by bulk async copies in the plane transforms, cuFFT instead of the x-line / plane kernels."""
    from gpu_common import GpuRun, OracleRun, rel_err
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cfg, pos, types, q = _system(20000, [32, 64, 64], [4.0, 5.0, 6.0], dtype, seed=31, coulomb=True)
    g = GpuRun(cfg, pos, types, charges=q)
    o = OracleRun(cfg, pos, types, charges=q)
    assert rel_err(g.forces(), o.force) < TOL[dtype]
    assert rel_err(g.eforces(), o.elec_forces) < TOL[dtype]
    for t in range(cfg.n_types):
        for d in range(3):
            assert rel_err(g.force_mesh[t][d].value.cpu().numpy(), o.st.force_mesh[t][d]) < TOL[dtype]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mesh", [[32, 32, 32], [24, 20, 28]])
def test_second_pme_call_on_another_particle_set(dtype, mesh):
    """The peptide-dipole call of main.py:1060-1095: ``update_field_force_q`` with the reconstructed dipole charges /
    positions (another particle set, 4 point charges per torsion) and meshes from ``pm.create``.  Forces on the
    dipoles equal the oracle's PME forces for that set, and the charge density / potential / electrostatic energy
    of the real charges are unchanged afterwards (the reference keeps the two in separate pmesh fields)."""
    from gpu_common import GpuRun, OracleRun, rel_err
    from hymd_b200 import field as F
    cfg, pos, types, q = _system(9000, mesh, [4.0, 5.0, 6.0], dtype, seed=71, coulomb=True)
    g = GpuRun(cfg, pos, types, charges=q)
    vel = torch.zeros_like(g.pos)
    e_before = g.energies(vel)
    psi_before = g.psi.value.cpu().numpy().copy()
    rng = np.random.default_rng(5)
    n_tors = 37
    dip_pos = (rng.uniform(0, 1, size=(4 * n_tors, 3)) * np.asarray(cfg.box_size)).astype(dtype)
    dip_q = np.tile(np.array([0.25, -0.25, 0.25, -0.25]), n_tors).astype(dtype)
    pm = g.pm
    phi_d, psi_d = pm.create("real", value=0.0), pm.create("real", value=0.0)
    phi_df, psi_df = pm.create("complex", value=0.0), pm.create("complex", value=0.0)
    field_df = [pm.create("complex", value=0.0) for _ in range(3)]
    field_d = [pm.create("real", value=0.0) for _ in range(3)]
    for as_numpy in (False, True):
        if as_numpy:
            dq, dp, df = dip_q, dip_pos, np.zeros((4 * n_tors, 3), dtype=dtype)
        else:
            dq = torch.as_tensor(dip_q, device="cuda")
            dp = torch.as_tensor(dip_pos, device="cuda")
            df = torch.zeros((4 * n_tors, 3), dtype=dq.dtype, device="cuda")
        F.update_field_force_q(dq, phi_d, phi_df, psi_d, psi_df, field_df, field_d, df, pm.decompose(dp),
                               g.h, pm, dp, cfg)
        got = df if as_numpy else df.cpu().numpy()
        o = OracleRun(cfg, dip_pos, np.zeros(4 * n_tors, dtype=np.int32), charges=dip_q)
        assert rel_err(got, o.elec_forces) < TOL[dtype]
    torch.cuda.synchronize()
    e_after = g.energies(vel)
    assert e_after == e_before
    assert np.array_equal(psi_before, g.psi.value.cpu().numpy())
    # ... and the real charges can be cycled again afterwards
    F.update_field_force_q(g.q, g.phi_q, g.phi_q_fourier, g.psi, None, None, g.elec_field, g.elec_forces,
                           pm.decompose(None), g.h, pm, g.pos, cfg)
    o2 = OracleRun(cfg, pos, types, charges=q)
    assert rel_err(g.eforces(), o2.elec_forces) < TOL[dtype]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_peptide_dipole_flow_matches_oracle(dtype):
    """main.py:530-552 + 1060-1095 end to end on the device: dihedral forces of a peptide backbone with
    ``dipole_flag = 1`` leave the reconstructed dipole charges' positions and the transfer matrices; the PME cycle on
    those 4 n_tors point charges (second context, the real charges' state untouched) gives the forces on them;
    ``dipole_forces_redistribution`` carries these back to the backbone beads.  Compared with the same chain of
    oracle restatements."""
    from gpu_common import GpuRun, OracleRun, rel_err
    from hymd_b200 import field as F
    from hymd_b200 import force as FO
    from oracle import bonded_oracle as bo
    cfg, pos, types, q = _system(6000, [32, 32, 32], [4.0, 5.0, 6.0], dtype, seed=91, coulomb=True)
    g = GpuRun(cfg, pos, types, charges=q)
    rng = np.random.default_rng(17)
    box = np.asarray(cfg.box_size, dtype=np.float64)
    # three backbones of 30 beads among the first particles: consecutive beads 0.3 nm apart
    n_bb, L = 3, 30
    a = []
    for c in range(n_bb):
        start = c * L
        chain = np.mod(np.cumsum(rng.normal(scale=0.2, size=(L, 3)), axis=0) + rng.uniform(0, 1, 3) * box, box)
        pos[start:start + L] = chain.astype(dtype)
        a += list(range(start, start + L - 3))
    a = np.array(a)
    n_tors = len(a)
    idx = (a, a + 1, a + 2, a + 3)
    coeff = np.zeros((n_tors, 6, 5))
    coeff[:, 0] = rng.uniform(0.5, 2.0, size=(n_tors, 5))
    coeff[:, 4] = rng.uniform(20, 60, size=(n_tors, 5))
    dt = np.ones(n_tors, dtype=int)
    last = np.zeros(n_tors, dtype=int)
    last[L - 4::L - 3] = 1
    charges_d = np.array([[0.25, -0.25, 0.25, -0.25] if l else [0.25, -0.25, 0.0, 0.0] for l in last],
                         dtype=dtype).reshape(-1)                      # main.py:475-485
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    pos_d = torch.as_tensor(pos, device="cuda")
    f_dih = torch.empty_like(pos_d)
    dip = torch.zeros((n_tors, 4, 3), dtype=tdt, device="cuda")
    tm = torch.zeros((n_tors, 6, 3, 3), dtype=tdt, device="cuda")
    FO.compute_dihedral_forces(f_dih, pos_d, dip, tm, box, *idx, coeff, dt, last, 1)
    pm = g.pm
    dip_pos = dip.reshape(4 * n_tors, 3)
    f_dip = torch.zeros_like(dip_pos)
    meshes = [pm.create("real"), pm.create("complex"), pm.create("real"), pm.create("complex"),
              [pm.create("complex") for _ in range(3)], [pm.create("real") for _ in range(3)]]
    F.update_field_force_q(torch.as_tensor(charges_d, device="cuda"), *meshes, f_dip, pm.decompose(dip_pos), g.h, pm,
                           dip_pos, cfg)
    f_beads = torch.zeros_like(pos_d)
    FO.dipole_forces_redistribution(f_beads, f_dip.reshape(n_tors, 4, 3), tm, *idx, dt, last, coeff=coeff)
    # oracle chain
    fo_, eo, dipo, tmo = bo.compute_dihedral_forces(pos, box, *idx, coeff, dt, last, dipole_flag=1, full=True)
    o = OracleRun(cfg, dipo.reshape(-1, 3), np.zeros(4 * n_tors, dtype=np.int32), charges=charges_d)
    want = bo.dipole_forces_redistribution(len(pos), o.elec_forces.reshape(n_tors, 4, 3), tmo, *idx, dt, last)
    tol = TOL[dtype]
    assert rel_err(f_dih.cpu().numpy(), fo_) < (1e-6 if dtype == np.float32 else 1e-10)
    assert rel_err(f_dip.cpu().numpy(), o.elec_forces) < 10 * tol     # dipole positions differ by one rounding in fp32
    assert rel_err(f_beads.cpu().numpy(), want) < 10 * tol


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mesh,n", [([32, 64, 64], 60000), ([24, 20, 28], 9000), ([16, 16, 80], 200000), ([9, 12, 10], 700)])
def test_paint_kernels_are_bitwise_identical(dtype, mesh, n, monkeypatch):
    """The flattened-list paint (default) and the row-walker paint (HYMD_B200_PAINT=rows: a lane per cell row,
    z-major accumulators, bulk-copy staged records) accumulate the same fixed-point contributions: densities and the
    charge density are equal bit for bit, and equal to the oracle within the tolerance.  The 16 x 16 x 80 system
    (~10 particles per cell) overflows the staging buffer: the tail of the runs is read from global memory."""
    from gpu_common import GpuRun, OracleRun, rel_err
    cfg, pos, types, q = _system(n, mesh, [4.0, 5.0, 6.0], dtype, seed=57, coulomb=True)
    a = GpuRun(cfg, pos, types, charges=q)
    phi_a = [a.phi[t].value.cpu().numpy().copy() for t in range(cfg.n_types)]
    rho_a = a.phi_q.value.cpu().numpy().copy()
    monkeypatch.setenv("HYMD_B200_PAINT", "rows")
    b = GpuRun(cfg, pos, types, charges=q)
    for t in range(cfg.n_types):
        assert np.array_equal(phi_a[t], b.phi[t].value.cpu().numpy())
    assert np.array_equal(rho_a, b.phi_q.value.cpu().numpy())
    o = OracleRun(cfg, pos, types, charges=q)
    for t in range(cfg.n_types):
        assert rel_err(phi_a[t], o.st.phi[t]) < TOL[dtype]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mesh,coulomb", [([24, 20, 28], True), ([32, 32, 32], False), ([9, 12, 10], True)])
def test_laplacian_and_pressure_match_oracle(dtype, mesh, coulomb):
    """comp_laplacian (field.py:406-425) and the 18 pressure contributions of comp_pressure
    (pressure.py:84-200) against the oracle's restatement."""
    from gpu_common import GpuRun, OracleRun, rel_err
    from hymd_b200 import field as F
    from hymd_b200.pressure import comp_pressure
    from oracle import field_oracle as fo
    cfg, pos, types, q = _system(9000, mesh, [4.0, 5.0, 6.0], dtype, seed=41, coulomb=coulomb)
    if coulomb:
        cfg.type_charges = [0.7, -0.4, 0.1]
        cfg.self_energy = 12.5
    rng = np.random.default_rng(9)
    vel = rng.normal(size=pos.shape).astype(dtype)
    bond_pr, angle_pr = np.array([1.0, -2.0, 0.5]), np.array([0.25, 0.5, -0.75])
    g = GpuRun(cfg, pos, types, charges=q, compute_potential=True)
    o = OracleRun(cfg, pos, types, charges=q)
    o.cfg.type_charges, o.cfg.self_energy = cfg.type_charges, getattr(cfg, "self_energy", 0.0)
    F.comp_laplacian(g.phi_fourier, g.phi_transfer, g.phi_laplacian, g.h, cfg)
    lap_ref = fo.comp_laplacian(o.st, o.cfg)
    scale = max(np.abs(lap_ref[t][d]).max() for t in range(cfg.n_types) for d in range(3))
    for t in range(cfg.n_types):
        for d in range(3):
            got = g.phi_laplacian[t][d].value.cpu().numpy()
            assert np.abs(got - lap_ref[t][d]).max() < TOL[dtype] * scale
    got = comp_pressure(g.phi, g.phi_q, g.psi, g.h, vel, cfg, g.phi_fourier, g.phi_laplacian,
                        g.phi_transfer, g.pos, bond_pr, angle_pr)
    want = fo.comp_pressure(o.st, o.h, vel, o.cfg, bond_pr, angle_pr)
    assert got.shape == (18,)
    # every term against the largest term (the total mixes terms of both signs)
    ref = np.abs(want).max()
    assert np.abs(got - want).max() < 10 * TOL[dtype] * ref, (got, want)


def test_positions_moved_behind_torchs_back_are_rebinned():
    """hymd_md_kick_drift writes positions through a raw pointer (torch's version counter
    does not move).  A read-out / PME call on the moved positions WITHOUT a preceding update_field must
    re-bin them (bins are keyed on pointer + version + the library's write epoch), not reuse stale bins."""
    from gpu_common import GpuRun, rel_err
    from hymd_b200 import field as F
    from hymd_b200.md import kick_drift
    cfg, pos, types, q = _system(5000, [16, 16, 16], [4.0, 5.0, 6.0], np.float64, seed=5, coulomb=True)
    g = GpuRun(cfg, pos, types, charges=q)
    vel = torch.full_like(g.pos, 3.0)
    kick_drift(vel, g.pos, [torch.zeros_like(g.pos)], 72.0, 0.0, 0.1, box=cfg.box_size)   # x += 0.3 nm, wrapped
    ef = torch.zeros_like(g.elec_forces)
    F.update_field_force_q(g.q, g.phi_q, g.phi_q_fourier, g.psi, None, None, g.elec_field, ef,
                           g.pm.decompose(None), g.h, g.pm, g.pos, cfg)
    fresh = GpuRun(cfg, g.pos.cpu().numpy(), types, charges=q)
    assert rel_err(ef.cpu().numpy(), fresh.eforces()) < 1e-12
    assert rel_err(ef.cpu().numpy(), g.eforces()) > 1e-3          # the particles really moved


def test_numpy_positions_mutated_in_place_are_rebinned():
    """numpy arrays carry no version counter; an in-place update between
    update_field and a later call is detected (strided-sample hash) and re-binned."""
    from gpu_common import GpuRun, rel_err
    from hymd_b200 import field as F
    cfg, pos, types, q = _system(5000, [16, 16, 16], [4.0, 5.0, 6.0], np.float64, seed=6, coulomb=True)
    g = GpuRun(cfg, pos, types, charges=q, as_numpy=True)
    before = g.eforces().copy()
    pos += 0.3
    np.mod(pos, np.asarray(cfg.box_size, dtype=pos.dtype)[None, :], out=pos)
    ef = np.zeros_like(before)
    F.update_field_force_q(g.q, g.phi_q, g.phi_q_fourier, g.psi, None, None, g.elec_field, ef,
                           g.pm.decompose(None), g.h, g.pm, pos, cfg)
    fresh = GpuRun(cfg, pos.copy(), types, charges=q, as_numpy=True)
    assert rel_err(ef, fresh.eforces()) < 1e-12
    assert rel_err(ef, before) > 1e-3


@pytest.mark.parametrize("coulomb", [False, True])
def test_tensor_memory_plane_transforms(coulomb, monkeypatch):
    """256 x 256 fp32 planes: the (y,z) transforms exchange their two phases through tensor memory
    (planefft.cu, plane_*_tmem_kernel).  Every output kind they produce -- force meshes (derive mode, ghost
    layout), filtered densities, potentials, psi, electric field, Laplacians -- against the oracle, and the
    forces against the L2-scratch kernels (HYMD_B200_PLANE_TMEM=0) they replace."""
    from gpu_common import GpuRun, OracleRun, rel_err
    cfg, pos, types, q = _system(40000, [16, 256, 256], [3.0, 9.0, 9.5], np.float32, seed=21, coulomb=coulomb)
    g = GpuRun(cfg, pos, types, charges=q, compute_potential=True)
    o = OracleRun(cfg, pos, types, charges=q)
    tol = TOL[np.float32]
    assert rel_err(g.forces(), o.force) < tol
    for t in range(cfg.n_types):
        assert rel_err(g.phi[t].value.cpu().numpy(), o.st.phi[t]) < tol
        scale = np.abs(o.st.v_ext[t]).max() + 1.0 / cfg.kappa
        assert np.abs(g.v_ext[t].value.cpu().numpy() - o.st.v_ext[t]).max() / scale < tol
        for d in range(3):
            assert rel_err(g.force_mesh[t][d].value.cpu().numpy(), o.st.force_mesh[t][d]) < tol
    from oracle import field_oracle as fo
    fo.comp_laplacian(o.st, o.cfg, workers=-1)
    lscale = max(np.abs(o.st.phi_laplacian[t][d]).max() for t in range(cfg.n_types) for d in range(3))
    for t in range(cfg.n_types):
        for d in range(3):
            got = g.phi_laplacian[t][d].value.cpu().numpy()
            assert np.abs(got - o.st.phi_laplacian[t][d]).max() / lscale < tol
    if coulomb:
        assert rel_err(g.eforces(), o.elec_forces) < tol
        assert rel_err(g.psi.value.cpu().numpy(), o.st.psi) < tol
        for d in range(3):
            assert rel_err(g.elec_field[d].value.cpu().numpy(), o.st.elec_field[d]) < tol
    monkeypatch.setenv("HYMD_B200_PLANE_TMEM", "0")
    h = GpuRun(cfg, pos, types, charges=q)
    assert rel_err(g.forces(), h.forces()) < 2e-6

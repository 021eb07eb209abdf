"""The slab-sharded cycle on ONE GPU: P ranks as P threads of the test process ("virtual slabs",
hymd_local_group_id), each with its own context (slab r of the mesh) and CUDA stream.

Everything the multi-GPU path does runs here unchanged -- the transposes and halos stored into the peers'
buffers, the per-step routing of particles that are not on the rank owning their slab (pmesh's
pm.decompose / Layout.exchange, main.py:977-980, field.py:574, 200), the slab-sharded energies -- only the
transport differs (plain device pointers instead of CUDA IPC mappings over NVLink, host barriers instead of
flag barriers in peer memory).  A single-GPU driver box therefore checks the sharded data flow against the
oracle at the north-star tolerances; tests/test_mgpu.py repeats it over real NVLink when the box has
several GPUs."""
import numpy as np
import pytest
import torch

from conftest import make_config

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

CHI3 = [("A", "B", 9.6754032616815161), ("A", "C", -13.2596290315913623),
        ("B", "C", 0.3852001771213374)]
TOL = {np.float32: 1e-5, np.float64: 1e-10}


def _system(n, mesh, dtype, pme, seed=11):
    rng = np.random.default_rng(seed)
    box = np.asarray([4.0, 5.0, 6.0], dtype=np.float32)
    pos = (rng.uniform(0, 1, size=(n, 3)) * box).astype(dtype)
    pos = np.minimum(pos, np.nextafter(box.astype(dtype), 0).astype(dtype))
    names = ["ABC"[i % 3] for i in range(n)]
    cfg = make_config(names, n, mesh, box, chi=CHI3, dtype=dtype,
                      coulombtype="PIC_Spectral" if pme else None,
                      dielectric_const=80.0 if pme else None)
    types = np.array([cfg.name_to_type_map[t] for t in names], dtype=np.int32)
    q = None
    if pme:
        q = rng.choice([-1.0, 0.0, 1.0], size=n)
        q -= q.mean()
        q = q.astype(dtype)
    return cfg, pos, types, q


def _ownership(pos, mesh, box, P, mode):
    n = len(pos)
    cell = np.floor(pos[:, 0].astype(np.float64) * mesh[0] / float(box[0])).astype(np.int64) % mesh[0]
    home = cell // (mesh[0] // P)
    if mode == "home":
        return home
    if mode == "scattered":              # arbitrary share: most particles are guests somewhere else
        return np.arange(n) % P
    if mode == "drifted":                # at home except a band next to every slab face (MD drift)
        owner = home.copy()
        frac = pos[:, 0].astype(np.float64) * mesh[0] / float(box[0]) / (mesh[0] // P)
        near = np.abs(frac - np.round(frac)) < 0.08
        owner[near] = (home[near] + 1) % P
        return owner
    if mode == "one_empty":              # rank P-1 holds nothing and still serves its slab
        return home % (P - 1) if P > 1 else home
    raise ValueError(mode)


def _cycle(rank, cfg, pos, types, q, owner, tdt, steps, results):
    from hymd_b200 import field as F
    from hymd_b200.hamiltonian import get_hamiltonian
    idx = np.nonzero(owner == rank)[0]
    ham = get_hamiltonian(cfg)
    pm, fl, ecl, cl = F.initialize_pm(None, cfg)
    phi, phi_fourier, force_mesh, v_ext_fourier, v_ext, phi_transfer, phi_laplacian = fl
    phi_q, phi_q_fourier, psi, elec_field = ecl
    dev = pm.device
    pos_d = torch.as_tensor(np.ascontiguousarray(pos[idx]), dtype=tdt, device=dev)
    typ_d = torch.as_tensor(types[idx], device=dev)
    q_d = None if q is None else torch.as_tensor(q[idx], dtype=tdt, device=dev)
    layouts = [pm.decompose(None) for _ in range(cfg.n_types)]
    force_d = torch.zeros((len(idx), 3), dtype=tdt, device=dev)
    eforce_d = torch.zeros((len(idx), 3), dtype=tdt, device=dev) if q is not None else None
    out = []
    box = torch.as_tensor(np.asarray(cfg.box_size), dtype=tdt, device=dev)
    for step in range(steps):
        if step:                          # move everybody a little: the next step re-uses the bin order
            pos_d = torch.remainder(pos_d + 0.013 * step, box)
            pos_d = torch.where(pos_d >= box, torch.zeros_like(pos_d), pos_d).contiguous()
        F.update_field(phi, phi_laplacian, phi_transfer, layouts, force_mesh, ham, pm, pos_d, typ_d,
                       cfg, v_ext, phi_fourier, v_ext_fourier, cfg.m, compute_potential=(step == steps - 1))
        F.compute_field_force(layouts, pos_d, force_mesh, force_d, typ_d, cfg.n_types)
        if q is not None:
            F.update_field_force_q(q_d, phi_q, phi_q_fourier, psi, None, None, elec_field, eforce_d,
                                   pm.decompose(None), ham, pm, pos_d, cfg)
    vel = torch.zeros_like(pos_d)
    energies = F.compute_field_and_kinetic_energy(phi, phi_q, psi, vel, ham, pos_d, typ_d, v_ext, cfg, layouts)
    pm.check()
    st = pm.status()
    cost = layouts[0].get_exchange_cost()           # collective (pmesh Layout.get_exchange_cost, main.py:1304-1312)
    results[rank] = {
        "exchange_cost": cost,
        "idx": idx, "pos": pos_d.cpu().numpy(), "force": force_d.cpu().numpy(),
        "eforce": None if eforce_d is None else eforce_d.cpu().numpy(),
        "phi": [p.value.cpu().numpy() for p in phi],
        "fmesh": [[force_mesh[t][d].value.cpu().numpy() for d in range(3)] for t in range(cfg.n_types)],
        "v_ext": [v.value.cpu().numpy() for v in v_ext],
        "psi": None if q is None else psi.value.cpu().numpy(),
        "energies": energies, "away": st["out_of_slab"], "paths": pm.paths(),
    }
    pm.close()
    return True


def _run_and_compare(P, mesh, n, dtype, pme, mode, steps=1, env=None, monkeypatch=None):
    from gpu_common import OracleRun, rel_err
    from hymd_b200._world import VirtualRanks
    for k, v in (env or {}).items():
        monkeypatch.setenv(k, v)
    cfg, pos, types, q = _system(n, mesh, dtype, pme)
    owner = _ownership(pos, mesh, cfg.box_size, P, mode)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    results = [None] * P
    VirtualRanks(P).run(_cycle, cfg, pos, types, q, owner, tdt, steps, results)
    tol = TOL[dtype]
    gpos = np.zeros_like(pos)
    force = np.full((n, 3), np.nan)
    eforce = np.full((n, 3), np.nan)
    for r in results:
        gpos[r["idx"]] = r["pos"]
        force[r["idx"]] = r["force"]
        if pme:
            eforce[r["idx"]] = r["eforce"]
    o = OracleRun(cfg, gpos, types, charges=q)
    checks = {"force": rel_err(force, o.force)}
    if pme:
        checks["eforce"] = rel_err(eforce, o.elec_forces)
        checks["psi"] = rel_err(np.concatenate([r["psi"] for r in results], axis=0), o.st.psi)
    for t in range(cfg.n_types):
        checks[f"phi{t}"] = rel_err(np.concatenate([r["phi"][t] for r in results], axis=0), o.st.phi[t])
        scale = np.abs(o.st.v_ext[t]).max() + 1.0 / cfg.kappa
        checks[f"v_ext{t}"] = np.abs(np.concatenate([r["v_ext"][t] for r in results], axis=0) - o.st.v_ext[t]).max() / scale
        for d in range(3):
            checks[f"fmesh{t}{d}"] = rel_err(np.concatenate([r["fmesh"][t][d] for r in results], axis=0),
                                             o.st.force_mesh[t][d])
    e_o = o.energies(np.zeros_like(gpos))
    e_g = results[0]["energies"]
    checks["field_energy"] = abs(e_g[0] - e_o[0]) / max(abs(e_o[0]), 0.5 * n / cfg.kappa * 1e-2)
    if pme:
        checks["field_q_energy"] = abs(e_g[2] - e_o[2]) / max(abs(e_o[2]), 1e-300)
    bad = {k: v for k, v in checks.items() if not v < tol}
    assert not bad, f"P={P} mesh={mesh} mode={mode}: {bad}"
    # Layout.get_exchange_cost: entry r = the particles rank r holds outside its own slab (its guests elsewhere)
    cell = np.floor(gpos[:, 0].astype(np.float64) * mesh[0] / float(cfg.box_size[0])).astype(np.int64) % mesh[0]
    home = cell // (mesh[0] // P)
    want = np.array([np.count_nonzero((owner == r) & (home != r)) for r in range(P)])
    for r in results:
        assert np.array_equal(r["exchange_cost"], want), (r["exchange_cost"], want)
    return results


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("P,mesh", [(2, [32, 24, 40]), (4, [32, 24, 40]), (2, [32, 32, 32]), (4, [16, 64, 64]),
                                    (8, [64, 32, 32]), (2, [18, 12, 10]), (2, [16, 256, 256])])
def test_virtual_slabs_match_oracle(P, mesh, dtype):
    """Particles at home: transposes, halo reduce / fetch, slab-sharded energies (cuFFT path for the odd
    meshes, plane + x-line kernels for the power-of-two ones)."""
    _run_and_compare(P, mesh, 20000, dtype, True, "home")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("P,mode", [(2, "scattered"), (4, "scattered"), (4, "drifted"), (8, "drifted"),
                                    (3, "one_empty")])
def test_guests_are_routed_to_their_slab_and_back(P, mode, dtype):
    """Row a2 / e: no particle needs to sit on the rank that owns its cell.  `scattered` hands every rank an
    arbitrary 1/P share (most particles are guests), `drifted` moves a band at every slab face to the
    neighbour rank, `one_empty` leaves a rank without particles of its own; three steps, so the second and
    third re-use the previous bin order with a different guest population."""
    mesh = [24, 24, 40] if P == 3 else [32, 32, 32]
    res = _run_and_compare(P, mesh, 20000, dtype, True, mode, steps=3)
    assert sum(r["away"] for r in res) > 0


def test_guest_capacity_overflow_fails_loudly(monkeypatch):
    """More guests than the buffers hold: an error on every rank (never a silently dropped particle)."""
    from hymd_b200._lib import HymdError
    with pytest.raises(HymdError, match="GUEST_CAPACITY"):
        _run_and_compare(2, [32, 32, 32], 20000, np.float64, False, "scattered",
                         env={"HYMD_B200_GUEST_CAPACITY": "64"}, monkeypatch=monkeypatch)


@pytest.mark.parametrize("mode", ["blocked", "fused", "kernels"])
@pytest.mark.parametrize("P,mesh", [(4, [32, 32, 32]), (2, [16, 256, 256])])
def test_exchange_modes(P, mesh, mode, monkeypatch):
    """The three ways the FFT transposes cross NVLink (HYMD_B200_EXCHANGE, slabfft.cu): per-destination blocks
    moved by contiguous peer copies (default), remote stores from the plane r2c / x-line epilogues, pack /
    unpack push kernels -- same results."""
    _run_and_compare(P, mesh, 20000, np.float32, True, "drifted", steps=2,
                     env={"HYMD_B200_EXCHANGE": mode}, monkeypatch=monkeypatch)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("P,mesh,pme", [(2, [32, 32, 32], True), (4, [32, 64, 64], False), (2, [16, 256, 256], False),
                                        (8, [64, 64, 64], True)])
def test_pipelined_exchange_matches_oracle(P, mesh, pme, dtype, monkeypatch):
    """HYMD_B200_XPIPE=2 forces the pipelined blocked exchange at every size (slabfft.cu): the forward transform runs
    field by field and the inverse one potential row at a time, each piece crossing on a second stream while the
    plane kernel works on the next one, with one barrier per row; HYMD_B200_XPIPE=0 is the one-copy exchange."""
    if dtype == np.float64 and mesh[1] == 256:
        pytest.skip("tensor-memory plane kernels are fp32")
    _run_and_compare(P, mesh, 20000, dtype, pme, "drifted", steps=3,
                     env={"HYMD_B200_XPIPE": "2"}, monkeypatch=monkeypatch)          # pieces moved by the copy engines
    _run_and_compare(P, mesh, 20000, dtype, pme, "drifted", steps=3,
                     env={"HYMD_B200_XPIPE": "2", "HYMD_B200_XPIPE_COPY": "kernel"}, monkeypatch=monkeypatch)
    # two fields / potential rows per piece (what the planner picks when a slab has fewer planes than the GPU has SMs):
    # three types -> pieces {0, 1} and {2}
    _run_and_compare(P, mesh, 20000, dtype, pme, "drifted", steps=2,
                     env={"HYMD_B200_XPIPE": "2", "HYMD_B200_XPIPE_GROUP": "2", "HYMD_B200_XPIPE_COPY": "ce"}, monkeypatch=monkeypatch)
    _run_and_compare(P, mesh, 20000, dtype, pme, "drifted", steps=2,
                     env={"HYMD_B200_XPIPE": "0"}, monkeypatch=monkeypatch)


@pytest.mark.parametrize("P,mesh", [(2, [32, 64, 64]), (4, [16, 128, 128])])
def test_half_size_plane_ctas_on_slabs(P, mesh, monkeypatch):
    """HYMD_B200_PLANE_THREADS=256 (two 256-thread CTAs per SM for 64^2 / 128^2 planes, the default for large
    launches) with the blocked receive layout of the slab pipeline, pipelined exchange forced on."""
    _run_and_compare(P, mesh, 20000, np.float32, True, "drifted", steps=2,
                     env={"HYMD_B200_PLANE_THREADS": "256", "HYMD_B200_XPIPE": "2"}, monkeypatch=monkeypatch)


def _dipole_cycle(rank, cfg, pos, types, q, owner, dip_pos, dip_q, dip_owner, tdt, results):
    """Main PME cycle on the rank's particles, then the peptide-dipole call (main.py:1060-1095) on the rank's own
    dipole charges with pm.create meshes: a second context per rank, created collectively on first use."""
    from hymd_b200 import field as F
    from hymd_b200.hamiltonian import get_hamiltonian
    idx = np.nonzero(owner == rank)[0]
    didx = np.nonzero(dip_owner == rank)[0]
    ham = get_hamiltonian(cfg)
    pm, fl, ecl, cl = F.initialize_pm(None, cfg)
    phi, phi_fourier, force_mesh, v_ext_fourier, v_ext, phi_transfer, phi_laplacian = fl
    phi_q, phi_q_fourier, psi, elec_field = ecl
    dev = pm.device
    pos_d = torch.as_tensor(np.ascontiguousarray(pos[idx]), dtype=tdt, device=dev)
    typ_d = torch.as_tensor(types[idx], device=dev)
    q_d = torch.as_tensor(q[idx], dtype=tdt, device=dev)
    layouts = [pm.decompose(None) for _ in range(cfg.n_types)]
    force_d = torch.zeros((len(idx), 3), dtype=tdt, device=dev)
    eforce_d = torch.zeros((len(idx), 3), dtype=tdt, device=dev)
    F.update_field(phi, phi_laplacian, phi_transfer, layouts, force_mesh, ham, pm, pos_d, typ_d, cfg, v_ext,
                   phi_fourier, v_ext_fourier, cfg.m)
    F.compute_field_force(layouts, pos_d, force_mesh, force_d, typ_d, cfg.n_types)
    F.update_field_force_q(q_d, phi_q, phi_q_fourier, psi, None, None, elec_field, eforce_d, pm.decompose(None), ham,
                           pm, pos_d, cfg)
    e0 = F.compute_field_and_kinetic_energy(phi, phi_q, psi, torch.zeros_like(pos_d), ham, pos_d, typ_d, v_ext, cfg, layouts)
    dp = torch.as_tensor(np.ascontiguousarray(dip_pos[didx]), dtype=tdt, device=dev)
    dq = torch.as_tensor(dip_q[didx], dtype=tdt, device=dev)
    df = torch.zeros((len(didx), 3), dtype=tdt, device=dev)
    F.update_field_force_q(dq, pm.create("real"), pm.create("complex"), pm.create("real"), pm.create("complex"),
                           [pm.create("complex") for _ in range(3)], [pm.create("real") for _ in range(3)], df,
                           pm.decompose(dp), ham, pm, dp, cfg)
    e1 = F.compute_field_and_kinetic_energy(phi, phi_q, psi, torch.zeros_like(pos_d), ham, pos_d, typ_d, v_ext, cfg, layouts)
    pm.check()
    results[rank] = {"didx": didx, "dforce": df.cpu().numpy(), "e0": e0, "e1": e1, "idx": idx,
                     "eforce": eforce_d.cpu().numpy()}
    pm.close()
    return True


@pytest.mark.parametrize("P", [2, 4])
def test_peptide_dipole_call_on_slabs(P):
    """The second PME call on another particle set, sharded: every rank holds its own dipole charges (they are not
    domain-decomposed in the reference either, main.py:469-471), the second context routes them like any particle;
    forces equal the single-rank oracle and the real charges' energies are untouched."""
    from gpu_common import OracleRun, rel_err
    from hymd_b200._world import VirtualRanks
    dtype, mesh = np.float64, [32, 32, 32]
    cfg, pos, types, q = _system(12000, mesh, dtype, True)
    owner = _ownership(pos, mesh, cfg.box_size, P, "drifted")
    rng = np.random.default_rng(4)
    nd = 4 * 53
    dip_pos = (rng.uniform(0, 1, size=(nd, 3)) * np.asarray(cfg.box_size)).astype(dtype)
    dip_q = np.tile(np.array([0.25, -0.25, 0.25, -0.25]), nd // 4).astype(dtype)
    dip_owner = np.arange(nd) // 4 % P                 # whole torsions on arbitrary ranks
    results = [None] * P
    VirtualRanks(P).run(_dipole_cycle, cfg, pos, types, q, owner, dip_pos, dip_q, dip_owner, torch.float64, results)
    df = np.zeros((nd, 3))
    ef = np.zeros((len(pos), 3))
    for r in results:
        df[r["didx"]] = r["dforce"]
        ef[r["idx"]] = r["eforce"]
        assert r["e0"] == r["e1"]
    o = OracleRun(cfg, dip_pos, np.zeros(nd, dtype=np.int32), charges=dip_q)
    assert rel_err(df, o.elec_forces) < TOL[dtype]
    o2 = OracleRun(cfg, pos, types, charges=q)
    assert rel_err(ef, o2.elec_forces) < TOL[dtype]

"""GPU parity of row f3 (general-Poisson-equation electrostatics, hymd_gpe_cycle) against
oracle/gpe_oracle.py, which is pinned on the reference's own update_field_force_q_GPE
(tests/test_oracle_gpe.py).  First run on a B200 by the round-1 driver (GPUTEST_r01: 11 XPASS)."""
import numpy as np
import pytest
import torch

from conftest import make_config

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]


@pytest.mark.parametrize("conv", ["csum", "euclidean_norm"])
def test_gpe_convergence_measures_match_oracle(monkeypatch, conv):
    """csum / euclidean_norm (main.py:141-152) stop the polarisation loop in the same iteration as the oracle."""
    test_gpe_matches_oracle(monkeypatch, np.float64, 1e-8, [16, 16, 16], conv=conv)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-8), (np.float32, 2e-4)])
@pytest.mark.parametrize("mesh", [[16, 16, 16], [10, 12, 8], [9, 8, 11]])
def test_gpe_matches_oracle(monkeypatch, dtype, tol, mesh, conv=None):
    from gpu_common import rel_err
    from hymd_b200 import field as F
    from hymd_b200.hamiltonian import get_hamiltonian
    from oracle import field_oracle as fo
    from oracle import gpe_oracle as go
    from oracle.hamiltonian_oracle import OracleHamiltonian
    rng = np.random.default_rng(61)
    n, box = 3000, np.array([3.5, 4.0, 3.0], dtype=np.float32)
    names = [("A", "B", "W", "W")[i % 4] for i in range(n)]
    cfg = make_config(names, n, mesh, box, chi=[("A", "B", 15.0), ("A", "W", 25.0)], dtype=dtype,
                      coulombtype="PIC_Spectral_GPE")
    cfg.type_charges = [1.0, -1.0, 0.0]
    cfg.dielectric_type = [5.0, 10.0, 80.0]
    cfg.pol_mixing, cfg.conv_crit, cfg.convergence_type = 0.6, 1e-6, conv
    types_ = np.array([cfg.name_to_type_map[t] for t in names], dtype=np.int32)
    pos = (rng.random((n, 3)) * box).astype(dtype)
    pos = np.minimum(pos, np.nextafter(box.astype(dtype), 0).astype(dtype))
    q = np.asarray(cfg.type_charges)[types_].astype(dtype)
    # --- oracle (float64 on the same float32-valued inputs)
    import copy
    ocfg = copy.deepcopy(cfg)
    W = OracleHamiltonian(ocfg)
    st = fo.FieldState(ocfg, np.float64)
    fo.update_field(st, W, pos.astype(np.float64), types_, ocfg)
    gs = go.GpeState(mesh, ocfg.n_types)
    f_ref = go.update_field_force_q_GPE(gs, st.phi, types_, q.astype(np.float64), pos.astype(np.float64), W, ocfg,
                                        conv=conv or "max_diff")
    e_ref = go.compute_field_energy_q_GPE(gs, ocfg)
    # --- device
    ham = get_hamiltonian(cfg)
    pm, fl, ecl, cl = F.initialize_pm(None, cfg)
    phi, phi_fourier, force_mesh, v_ext_fourier, v_ext, phi_transfer, phi_laplacian = fl
    phi_q, phi_q_fourier, psi, elec_field = ecl
    (phi_eps, phi_eps_fourier, phi_eta, phi_eta_fourier, phi_pol, phi_pol_prev, elec_dot, elec_field_contrib,
     Vbar_elec, Vbar_elec_fourier, force_mesh_elec, force_mesh_elec_fourier) = cl
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    dpos = torch.tensor(pos, dtype=tdt, device="cuda")
    dtyp = torch.tensor(types_, device="cuda")
    dq = torch.tensor(q, dtype=tdt, device="cuda")
    layouts = [pm.decompose(None) for _ in range(cfg.n_types)]
    F.update_field(phi, phi_laplacian, phi_transfer, layouts, force_mesh, ham, pm, dpos, dtyp, cfg, v_ext,
                   phi_fourier, v_ext_fourier, cfg.m)
    force = torch.zeros((n, 3), dtype=tdt, device="cuda")
    F.compute_field_force(layouts, dpos, force_mesh, force, dtyp, cfg.n_types)
    elec_forces = torch.zeros((n, 3), dtype=tdt, device="cuda")
    Vbar, eps, dot = F.update_field_force_q_GPE(
        None, phi, dtyp, dq, phi_q, phi_q_fourier, phi_eps, phi_eps_fourier, phi_eta, phi_eta_fourier,
        phi_pol_prev, phi_pol, elec_field, elec_forces, elec_field_contrib, psi, Vbar_elec, Vbar_elec_fourier,
        force_mesh_elec, force_mesh_elec_fourier, ham, pm.decompose(None), layouts, pm, dpos, cfg)
    energy = F.compute_field_energy_q_GPE(cfg, eps, 0.0, dot)
    torch.cuda.synchronize()
    assert pm.gpe_iterations == gs.iterations
    assert rel_err(eps.value.cpu().numpy(), gs.phi_eps) < tol
    assert rel_err(psi.value.cpu().numpy(), gs.psi) < tol
    assert rel_err(dot.value.cpu().numpy(), gs.elec_dot) < 10 * tol
    for t in range(cfg.n_types):
        assert rel_err(Vbar[t].value.cpu().numpy(), gs.Vbar_elec[t]) < 10 * tol
    assert rel_err(elec_forces.cpu().numpy(), f_ref) < 10 * tol
    assert energy == pytest.approx(e_ref, rel=10 * tol)
    # a second call starts the polarisation iteration from zero again (reference semantics) and, with
    # unchanged inputs, reproduces the first one bit for bit
    again = torch.zeros_like(elec_forces)
    F.update_field(phi, phi_laplacian, phi_transfer, layouts, force_mesh, ham, pm, dpos, dtyp, cfg, v_ext,
                   phi_fourier, v_ext_fourier, cfg.m)
    F.update_field_force_q_GPE(
        None, phi, dtyp, dq, phi_q, phi_q_fourier, phi_eps, phi_eps_fourier, phi_eta, phi_eta_fourier,
        phi_pol_prev, phi_pol, elec_field, again, elec_field_contrib, psi, Vbar_elec, Vbar_elec_fourier,
        force_mesh_elec, force_mesh_elec_fourier, ham, pm.decompose(None), layouts, pm, dpos, cfg)
    assert torch.equal(again, elec_forces)


def test_gpe_needs_no_opt_in(monkeypatch):
    """The HYMD_B200_ENABLE_GPE gate of round 1 is gone: the coulombtype works out of the box."""
    monkeypatch.delenv("HYMD_B200_ENABLE_GPE", raising=False)
    from hymd_b200 import field as F
    cfg = make_config(["A", "B"], 10, 8, [2.0, 2.0, 2.0], coulombtype="PIC_Spectral_GPE")
    pm, fl, ecl, cl = F.initialize_pm(None, cfg)
    assert len(cl) == 12


def _gpe_rank(rank, cfg, pos, types_, q, owner, tdt, results):
    from hymd_b200 import field as F
    from hymd_b200.hamiltonian import get_hamiltonian
    idx = np.nonzero(owner == rank)[0]
    ham = get_hamiltonian(cfg)
    pm, fl, ecl, cl = F.initialize_pm(None, cfg)
    phi, phi_fourier, force_mesh, v_ext_fourier, v_ext, phi_transfer, phi_laplacian = fl
    phi_q, phi_q_fourier, psi, elec_field = ecl
    (phi_eps, phi_eps_fourier, phi_eta, phi_eta_fourier, phi_pol, phi_pol_prev, elec_dot, elec_field_contrib,
     Vbar_elec, Vbar_elec_fourier, force_mesh_elec, force_mesh_elec_fourier) = cl
    dev = pm.device
    dpos = torch.as_tensor(np.ascontiguousarray(pos[idx]), dtype=tdt, device=dev)
    dtyp = torch.as_tensor(types_[idx], device=dev)
    dq = torch.as_tensor(q[idx], dtype=tdt, device=dev)
    layouts = [pm.decompose(None) for _ in range(cfg.n_types)]
    F.update_field(phi, phi_laplacian, phi_transfer, layouts, force_mesh, ham, pm, dpos, dtyp, cfg, v_ext,
                   phi_fourier, v_ext_fourier, cfg.m)
    elec_forces = torch.zeros((len(idx), 3), dtype=tdt, device=dev)
    Vbar, eps, dot = F.update_field_force_q_GPE(
        None, phi, dtyp, dq, phi_q, phi_q_fourier, phi_eps, phi_eps_fourier, phi_eta, phi_eta_fourier,
        phi_pol_prev, phi_pol, elec_field, elec_forces, elec_field_contrib, psi, Vbar_elec, Vbar_elec_fourier,
        force_mesh_elec, force_mesh_elec_fourier, ham, pm.decompose(None), layouts, pm, dpos, cfg)
    energy = F.compute_field_energy_q_GPE(cfg, eps, 0.0, dot)
    pm.check()
    results[rank] = {"idx": idx, "f": elec_forces.cpu().numpy(), "energy": energy, "iters": pm.gpe_iterations,
                     "eps": eps.value.cpu().numpy(), "psi": psi.value.cpu().numpy()}
    pm.close()
    return True


@pytest.mark.parametrize("conv", [None, "csum", "euclidean_norm"])
@pytest.mark.parametrize("P,mesh", [(2, [16, 16, 16]), (4, [32, 16, 16])])
def test_gpe_on_slabs_matches_oracle(P, mesh, conv):
    """GPE electrostatics sharded over P slabs (virtual ranks on one GPU): transforms through the slab pipeline,
    pointwise kernels on the owned planes, the convergence measure (max / sum / sum of squares) combined over the
    ranks on the device, guests read out on the owner of their cell and returned -- same iteration count, dielectric,
    potential, forces and energy as the single-rank oracle."""
    import copy
    from gpu_common import rel_err
    from hymd_b200._world import VirtualRanks
    from oracle import field_oracle as fo
    from oracle import gpe_oracle as go
    from oracle.hamiltonian_oracle import OracleHamiltonian
    dtype, tol = np.float64, 1e-8
    rng = np.random.default_rng(67)
    n, box = 4000, np.array([3.5, 4.0, 3.0], dtype=np.float32)
    names = [("A", "B", "W", "W")[i % 4] for i in range(n)]
    cfg = make_config(names, n, mesh, box, chi=[("A", "B", 15.0), ("A", "W", 25.0)], dtype=dtype,
                      coulombtype="PIC_Spectral_GPE")
    cfg.type_charges = [1.0, -1.0, 0.0]
    cfg.dielectric_type = [5.0, 10.0, 80.0]
    cfg.pol_mixing, cfg.conv_crit, cfg.convergence_type = 0.6, 1e-6, conv
    types_ = np.array([cfg.name_to_type_map[t] for t in names], dtype=np.int32)
    pos = (rng.random((n, 3)) * box).astype(dtype)
    q = np.asarray(cfg.type_charges)[types_].astype(dtype)
    ocfg = copy.deepcopy(cfg)
    W = OracleHamiltonian(ocfg)
    st = fo.FieldState(ocfg, np.float64)
    fo.update_field(st, W, pos, types_, ocfg)
    gs = go.GpeState(mesh, ocfg.n_types)
    f_ref = go.update_field_force_q_GPE(gs, st.phi, types_, q, pos, W, ocfg, conv=conv or "max_diff")
    e_ref = go.compute_field_energy_q_GPE(gs, ocfg)
    cell = np.floor(pos[:, 0] * mesh[0] / float(box[0])).astype(np.int64) % mesh[0]
    owner = (cell // (mesh[0] // P) + (np.arange(n) % 5 == 0)) % P          # a fifth of the particles are guests
    results = [None] * P
    VirtualRanks(P).run(_gpe_rank, cfg, pos, types_, q, owner, torch.float64, results)
    f = np.zeros((n, 3))
    for r in results:
        f[r["idx"]] = r["f"]
        assert r["iters"] == gs.iterations
        assert r["energy"] == pytest.approx(e_ref, rel=10 * tol)
    assert rel_err(np.concatenate([r["eps"] for r in results], axis=0), gs.phi_eps) < tol
    assert rel_err(np.concatenate([r["psi"] for r in results], axis=0), gs.psi) < tol
    assert rel_err(f, f_ref) < 10 * tol

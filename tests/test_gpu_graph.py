"""Graph replay of the per-step field update (hymd_update_cycle, csrc/graph.cu): a context that replays recorded
steps must return, bit for bit, what a context issuing the ordinary launches returns -- over moving positions,
alternating position buffers, numpy inputs, a box change, a particle-count change and with charges -- and both
must match the oracle.  The deterministic paint makes bitwise equality the right bar."""
import copy

import numpy as np
import pytest
import torch

from hymd_b200 import field as F
from hymd_b200.hamiltonian import get_hamiltonian
from hymd_b200.synthetic import make_system

pytestmark = pytest.mark.gpu


class Stepper:
    """initialize_pm once, then update_field + compute_field_force (+ PME) per call."""

    def __init__(self, cfg, n, graph, charges=False):
        self.cfg = cfg
        self.h = get_hamiltonian(cfg)
        self.pm, fl, ecl, _ = F.initialize_pm(None, cfg)
        self.pm.set_graph(graph)
        (self.phi, self.phi_fourier, self.force_mesh, self.v_ext_fourier, self.v_ext, self.phi_transfer,
         self.phi_laplacian) = fl
        self.phi_q, self.phi_q_fourier, self.psi, self.elec_field = ecl
        self.tdt = torch.float64 if np.dtype(cfg.dtype) == np.float64 else torch.float32
        self.layouts = [self.pm.decompose(None) for _ in range(cfg.n_types)]
        self.charges = charges

    def step(self, pos, types, q=None):
        n = len(pos)
        numpy_in = not isinstance(pos, torch.Tensor)
        force = np.zeros((n, 3), dtype=self.cfg.dtype) if numpy_in else torch.zeros((n, 3), dtype=self.tdt, device="cuda")
        F.update_field(self.phi, self.phi_laplacian, self.phi_transfer, self.layouts, self.force_mesh, self.h,
                       self.pm, pos, types, self.cfg, self.v_ext, self.phi_fourier, self.v_ext_fourier, self.cfg.m)
        F.compute_field_force(self.layouts, pos, self.force_mesh, force, types, self.cfg.n_types)
        out = [force if numpy_in else force.cpu().numpy()]
        if q is not None:
            ef = torch.zeros((n, 3), dtype=self.tdt, device="cuda")
            F.update_field_force_q(q, self.phi_q, self.phi_q_fourier, self.psi, None, None, self.elec_field, ef,
                                   self.pm.decompose(None), self.h, self.pm, pos, self.cfg)
            out.append(ef.cpu().numpy())
        torch.cuda.synchronize()
        return out


def frames_of(s, k, dtype):
    L = np.asarray(s.config.box_size, dtype=np.float64)
    out = []
    for i in range(k):
        f = np.mod(s.positions.astype(np.float64) + i * 0.25 * s.velocities.astype(np.float64), L).astype(dtype)
        f[f >= L.astype(dtype)] = 0
        out.append(np.ascontiguousarray(f))
    return out


@pytest.mark.parametrize("mesh", [24, 32, 64])           # cuFFT path / plane + x-line kernels
@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_replayed_steps_equal_ordinary_launches(mesh, dtype):
    from gpu_common import OracleRun, rel_err
    s = make_system("C1", dtype=dtype, mesh=mesh)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    fr = frames_of(s, 4, dtype)
    typ = torch.as_tensor(s.types.astype(np.int32), device="cuda")
    plain, graph = Stepper(s.config, len(fr[0]), False), Stepper(s.config, len(fr[0]), True)
    # (a) one position tensor updated in place, as an MD loop does
    pos = torch.as_tensor(fr[0], dtype=tdt, device="cuda")
    order = [0, 1, 2, 3, 2, 1, 0, 1, 2, 3]
    last = None
    for k in order:
        pos.copy_(torch.as_tensor(fr[k], dtype=tdt, device="cuda"))
        a, b = plain.step(pos, typ)[0], graph.step(pos, typ)[0]
        assert np.array_equal(a, b), f"frame {k}: replayed step differs from the ordinary launches"
        last = a
    st = graph.pm.graph_stats()
    assert st["recorded"] >= 2 and st["replayed"] >= 4, st
    assert plain.pm.graph_stats()["replayed"] == 0
    assert plain.pm.launch_count() == graph.pm.launch_count()
    # (b) a different tensor every step (the copy node's source is patched), then numpy inputs
    for k in [3, 0, 2, 1]:
        p = torch.as_tensor(fr[k], dtype=tdt, device="cuda")
        assert np.array_equal(plain.step(p, typ)[0], graph.step(p, typ)[0])
    replayed = graph.pm.graph_stats()["replayed"]
    assert replayed > st["replayed"]
    typ_n = s.types.astype(np.int32)
    for k in [1, 2, 3, 0, 1, 2]:
        assert np.array_equal(plain.step(fr[k], typ_n)[0], graph.step(fr[k], typ_n)[0])
    # (c) and the replayed result is the reference's
    tol = 1e-5 if dtype == np.float32 else 1e-10          # north-star tolerance
    o = OracleRun(s.config, fr[2].astype(np.float64), s.types, compute_potential=False)
    assert rel_err(graph.step(fr[2], typ_n)[0], o.force) < tol
    assert last is not None
    plain.pm.close()
    graph.pm.close()


def test_box_and_particle_count_changes_invalidate_recorded_steps():
    from gpu_common import OracleRun, rel_err
    dtype = np.float32
    s = make_system("C1", dtype=dtype, mesh=32)
    fr = frames_of(s, 3, dtype)
    typ = torch.as_tensor(s.types.astype(np.int32), device="cuda")
    plain, graph = Stepper(s.config, len(fr[0]), False), Stepper(s.config, len(fr[0]), True)
    pos = torch.as_tensor(fr[0], device="cuda")
    for k in [0, 1, 2, 1, 0, 1]:
        pos.copy_(torch.as_tensor(fr[k], device="cuda"))
        assert np.array_equal(plain.step(pos, typ)[0], graph.step(pos, typ)[0])
    assert graph.pm.graph_stats()["replayed"] >= 1
    # barostat-like box change (barostat.py:158-164): scaled box and coordinates
    box2 = (np.asarray(s.config.box_size, dtype=np.float64) * 1.03).astype(np.float32).astype(np.float64)
    cfg2 = copy.deepcopy(s.config)
    cfg2.box_size = box2.astype(np.float32)
    for st in (plain, graph):
        st.pm.set_box(box2)
        st.cfg = cfg2
        st.h = get_hamiltonian(cfg2)
    fr2 = [np.ascontiguousarray((f.astype(np.float64) * 1.03).astype(dtype)) for f in fr]
    for k in [0, 1, 2, 1, 0, 1, 2]:
        pos.copy_(torch.as_tensor(fr2[k], device="cuda"))
        a, b = plain.step(pos, typ)[0], graph.step(pos, typ)[0]
        assert np.array_equal(a, b), "after set_box"
    o = OracleRun(cfg2, fr2[2].astype(np.float64), s.types, compute_potential=False)
    assert rel_err(b, o.force) < 1e-5
    # fewer particles (a rank's count changes after domain_decomposition)
    n2 = len(fr[0]) - 777
    typ2 = typ[:n2].clone()
    pos2 = torch.empty((n2, 3), dtype=torch.float32, device="cuda")
    for k in [0, 1, 2, 1, 0, 1]:
        pos2.copy_(torch.as_tensor(fr2[k][:n2], device="cuda"))
        assert np.array_equal(plain.step(pos2, typ2)[0], graph.step(pos2, typ2)[0]), "after the count change"
    assert graph.pm.graph_stats()["recorded"] >= 5
    plain.pm.close()
    graph.pm.close()


def test_replay_with_charges_and_the_pme_call():
    dtype = np.float32
    s = make_system("C3", dtype=dtype, n=40_000, mesh=32)
    assert s.charges is not None
    fr = frames_of(s, 3, dtype)
    typ = torch.as_tensor(s.types.astype(np.int32), device="cuda")
    q = torch.as_tensor(s.charges, dtype=torch.float32, device="cuda")
    plain, graph = Stepper(s.config, len(fr[0]), False), Stepper(s.config, len(fr[0]), True)
    pos = torch.as_tensor(fr[0], device="cuda")
    for k in [0, 1, 2, 1, 0, 1, 2, 1]:
        pos.copy_(torch.as_tensor(fr[k], device="cuda"))
        a, b = plain.step(pos, typ, q), graph.step(pos, typ, q)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert graph.pm.graph_stats()["replayed"] >= 2
    plain.pm.close()
    graph.pm.close()


def test_timing_and_compute_potential_run_the_ordinary_launches():
    s = make_system("C1", dtype=np.float32, mesh=32)
    typ = torch.as_tensor(s.types.astype(np.int32), device="cuda")
    pos = torch.as_tensor(np.ascontiguousarray(s.positions), dtype=torch.float32, device="cuda")
    g = Stepper(s.config, len(pos), True)
    for _ in range(4):
        g.step(pos, typ)
    before = g.pm.graph_stats()
    assert before["replayed"] >= 1
    g.pm.set_timing(True)
    g.step(pos, typ)
    assert g.pm.graph_stats()["replayed"] == before["replayed"]
    assert "paint" in g.pm.timings()
    g.pm.set_timing(False)
    F.update_field(g.phi, g.phi_laplacian, g.phi_transfer, g.layouts, g.force_mesh, g.h, g.pm, pos, typ, g.cfg,
                   g.v_ext, g.phi_fourier, g.v_ext_fourier, g.cfg.m, compute_potential=True)
    assert g.pm.graph_stats()["replayed"] == before["replayed"]
    # (the by-product buffers allocated by that call changed the digest: one ordinary step per half of the record
    # double buffer, then the step is recorded again)
    for _ in range(4):
        g.step(pos, typ)
    assert g.pm.graph_stats()["replayed"] > before["replayed"]
    g.pm.close()

"""hymd_b200.barostat (Berendsen + SCR) against the reference's own barostat.py / barostat_scr.py.

tests/golden/barostat_golden.npz holds what the unmodified reference functions do to the box and the
positions when ``comp_pressure`` returns a prescribed vector (tests/golden/make_reference_golden.py);
the same vector is injected here, so this checks the scaling arithmetic, the prng call order, the
in-place semantics (including the reference's row-wise ``positions[:][0:2]`` scaling) and the
``change`` flag on the CPU.  The pressure itself is ``hymd_b200.pressure.comp_pressure`` (GPU parity in
tests/test_gpu_parity.py)."""
import os
import types

import numpy as np
import pytest

import hymd_b200.barostat as B

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "barostat_golden.npz"))


class FakePM:
    def __init__(self):
        self.boxes = []

    def set_box(self, box):
        self.boxes.append(np.array(box, dtype=np.float64))


@pytest.mark.parametrize("kind", ["berendsen", "scr"])
@pytest.mark.parametrize("fn", ["isotropic", "semiisotropic"])
@pytest.mark.parametrize("step,P_L,P_N", [(4, 1.0, 1.0), (4, 1.0, None), (5, 1.0, 1.0)])
def test_barostat_matches_reference(monkeypatch, kind, fn, step, P_L, P_N):
    monkeypatch.setattr(B, "comp_pressure", lambda *a, **k: G["barostat/pressure"].copy())
    cfg = types.SimpleNamespace(time_step=0.03, respa_inner=5, n_b=2, tau_p=1.5, target_temperature=323.0,
                                gas_constant=0.0083144621, box_size=np.array([5.0, 6.0, 7.0]),
                                target_pressure=types.SimpleNamespace(P_L=P_L, P_N=P_N))
    pos = G["barostat/pos0"].copy()
    pm = FakePM()
    stuff = (pm, "field_list", "elec", "coulomb")
    res, change = getattr(getattr(B, kind), fn)(None, stuff, None, None, None, None, pos, None, cfg, None, None,
                                                None, np.zeros(3), np.zeros(3), step, np.random.default_rng(99))
    pre = f"barostat/{kind}/{fn}/{step}_{P_L}_{P_N}"
    np.testing.assert_allclose(cfg.box_size, G[pre + "/box"], rtol=1e-15)
    np.testing.assert_allclose(pos, G[pre + "/pos"], rtol=1e-15)
    ref_change, ref_reinit = G[pre + "/change"]
    assert bool(change) == bool(ref_change)
    assert res is stuff                              # the handles the caller holds stay valid
    assert (len(pm.boxes) == 1) == bool(ref_reinit)  # context told about the new box exactly when the
    if pm.boxes:                                     # reference re-runs initialize_pm
        np.testing.assert_allclose(pm.boxes[0], G[pre + "/box"], rtol=1e-15)
    assert float(getattr(cfg, "surface_tension", 0.0) or 0.0) == pytest.approx(float(G[pre + "/surface_tension"]), rel=1e-14)


def test_torch_positions_scale_like_numpy(monkeypatch):
    import torch
    monkeypatch.setattr(B, "comp_pressure", lambda *a, **k: G["barostat/pressure"].copy())
    cfg = types.SimpleNamespace(time_step=0.03, respa_inner=5, n_b=2, tau_p=1.5, target_temperature=323.0,
                                gas_constant=0.0083144621, box_size=np.array([5.0, 6.0, 7.0]),
                                target_pressure=types.SimpleNamespace(P_L=1.0, P_N=1.0))
    pos = torch.tensor(G["barostat/pos0"])
    B.semiisotropic(None, (FakePM(),), None, None, None, None, pos, None, cfg, None, None, None, np.zeros(3),
                    np.zeros(3), 4, None)
    np.testing.assert_allclose(pos.numpy(), G["barostat/berendsen/semiisotropic/4_1.0_1.0/pos"], rtol=1e-15)

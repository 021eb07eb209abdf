"""Oracle Hamiltonian vs golden vectors produced by the reference's real hamiltonian.py
(tests/golden/make_golden.py, run in the build container)."""
import json
import os

import numpy as np
import pytest

from hymd_b200.config import Chi, Config
from oracle.hamiltonian_oracle import OracleHamiltonian

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "hamiltonian_golden.npz"))
with open(os.path.join(HERE, "golden", "hamiltonian_golden.json")) as fh:
    CASES = json.load(fh)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_hamiltonian_matches_reference(case):
    cfg = Config(mesh_size=16, sigma=case["sigma"], kappa=case["kappa"], box_size=case["box"],
                 hamiltonian=case["kind"], chi=[Chi(*c) for c in case["chi"]],
                 coulombtype=case.get("coulombtype"), dielectric_const=case.get("dielectric_const"),
                 self_energy=case.get("self_energy"))
    cfg.finalize(case["names"], n_particles=case["n"])
    if case.get("type_charges") is not None:
        cfg.type_charges = list(case["type_charges"])
    if not case.get("f32_params"):
        cfg.box_size = np.asarray(cfg.box_size, dtype=np.float64)
    W = OracleHamiltonian(cfg)
    pre = case["name"]
    rho0, a, vol = GOLD[pre + "/rho0_a_vol"]
    if case.get("f32_params"):
        # the reference's lambdas carry float32-precision rho0 and a ~7-digit `a`
        # (see make_golden.py); only a loose comparison is meaningful
        phi = list(GOLD[pre + "/phi"])
        scale = np.abs(GOLD[pre + "/w_0"]).max()
        np.testing.assert_allclose(W.w_0(phi), GOLD[pre + "/w_0"], rtol=0, atol=1e-5 * scale)
        for t in range(cfg.n_types):
            vs = np.abs(GOLD[pre + "/v_ext"][t]).max()
            np.testing.assert_allclose(W.v_ext[t](phi), GOLD[pre + "/v_ext"][t], rtol=0,
                                       atol=1e-5 * vs)
        return
    assert cfg.rho0 == pytest.approx(rho0, rel=1e-15)
    assert cfg.a == pytest.approx(a, rel=1e-15)
    assert cfg.simulation_volume == pytest.approx(vol, rel=1e-15)
    phi = list(GOLD[pre + "/phi"])
    k = [GOLD[pre + "/k0"], GOLD[pre + "/k1"], GOLD[pre + "/k2"]]
    np.testing.assert_allclose(W.H(k, GOLD[pre + "/v"]), GOLD[pre + "/H"], rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(W.w_0(phi), GOLD[pre + "/w_0"], rtol=1e-11, atol=1e-9)
    for t in range(cfg.n_types):
        np.testing.assert_allclose(W.v_ext[t](phi), GOLD[pre + "/v_ext"][t], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(W.w_elec([GOLD[pre + "/phi_q"], GOLD[pre + "/psi"]]),
                               GOLD[pre + "/w_elec"], rtol=1e-13, atol=1e-13)
    # pressure helpers (pressure.py:105-108)
    psi = GOLD[pre + "/psi"]
    for t in range(cfg.n_types):
        np.testing.assert_allclose(W.V_bar_0[t](phi), GOLD[pre + "/V_bar_0"][t], rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(W.V_bar[t]([phi, psi]), GOLD[pre + "/V_bar"][t], rtol=1e-11,
                                   atol=1e-9)

"""oracle/gpe_oracle.py against the reference's own ``update_field_force_q_GPE`` /
``compute_field_energy_q_GPE`` (hymd/field.py:964-1112, 706-760) executed over the pmesh stand-in
(tests/golden/make_reference_golden.py -> tests/golden/gpe_golden.npz).  Groundwork for SURVEY.md
section 8 row f3; the product raises NotImplementedError for this coulombtype."""
import os
import types

import numpy as np
import pytest

from conftest import make_config
from oracle import gpe_oracle as go
from oracle.hamiltonian_oracle import OracleHamiltonian

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpe_golden.npz"))

CASES = [
    dict(name="gpe3", names=["A", "B", "W"], n=900, mesh=[10, 12, 8], box=[3.5, 4.0, 3.0], sigma=0.5,
         chi=[("A", "B", 15.0), ("A", "W", 25.0)], type_charges=[1.0, -1.0, 0.0],
         dielectric_type=[5.0, 10.0, 80.0], pol_mixing=0.6, conv_crit=1e-6),
    dict(name="gpe2_odd", names=["A", "W"], n=700, mesh=[9, 8, 11], box=[3.0, 3.2, 3.4], sigma=0.45,
         chi=[("A", "W", 10.0)], type_charges=[0.5, -0.3333333333333333], dielectric_type=[20.0, 60.0],
         pol_mixing=0.5, conv_crit=1e-7),
]


def close(x, ref, tol=1e-10):
    ref = np.asarray(ref)
    np.testing.assert_allclose(np.asarray(x), ref, rtol=0, atol=tol * max(np.abs(ref).max(), 1e-300))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_gpe_oracle_reproduces_the_reference(case):
    pre = "gpe/" + case["name"]
    cfg = make_config(case["names"], case["n"], case["mesh"], case["box"], sigma=case["sigma"], chi=case["chi"],
                      coulombtype="PIC_Spectral_GPE")
    cfg.box_size = np.asarray(cfg.box_size, dtype=np.float64)
    cfg.type_charges = list(case["type_charges"])
    cfg.dielectric_type = list(case["dielectric_type"])
    cfg.pol_mixing, cfg.conv_crit = case["pol_mixing"], case["conv_crit"]
    W = OracleHamiltonian(cfg)
    st = go.GpeState(case["mesh"], cfg.n_types)
    f = go.update_field_force_q_GPE(st, list(G[pre + "/phi"]), G[pre + "/types"], G[pre + "/charges"],
                                    G[pre + "/pos"], W, cfg)
    assert 1 < st.iterations < 100            # the fixed-point iteration converged
    close(st.phi_eps, G[pre + "/phi_eps"], 1e-12)
    close(st.psi, G[pre + "/psi"])
    close(st.elec_dot, G[pre + "/elec_dot"])
    close(np.stack(st.Vbar_elec), G[pre + "/Vbar_elec"])
    close(f, G[pre + "/elec_forces"])
    assert go.compute_field_energy_q_GPE(st, cfg) == pytest.approx(float(G[pre + "/energy"]), rel=1e-10)

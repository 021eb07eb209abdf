"""oracle/field_oracle.py (the restatement) against the REAL hymd/field.py and hymd/pressure.py.

tests/golden/field_golden.npz holds what the reference's unmodified ``update_field``,
``compute_field_force``, ``update_field_force_q``, ``compute_field_and_kinetic_energy`` and
``comp_pressure`` produce when executed (in the build container, tests/golden/
make_reference_golden.py) on oracle/pmesh_standin.py, i.e. on the same restated pmesh primitives the
oracle uses.  Agreement here pins the oracle's operation sequence (where the filter is applied, the
-ik gradient, the Poisson divisor, the potential / pressure bookkeeping) on the reference's own code
path; the pmesh primitives themselves are pinned by tests/test_oracle_reference_kats.py and
tests/test_oracle_analytic.py."""
import os

import numpy as np
import pytest

from conftest import make_config
from oracle import field_oracle as fo
from oracle.hamiltonian_oracle import OracleHamiltonian

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "field_golden.npz"))

CASES = [
    dict(name="chi3_even", names=["A", "B", "C"], n=600, mesh=[12, 10, 8], box=[4.0, 3.5, 3.0],
         kind="DefaultWithChi", sigma=0.5, kappa=0.05,
         chi=[("A", "B", 20.0), ("A", "C", -5.0), ("B", "C", 10.0)]),
    dict(name="nochi_odd", names=["A", "B"], n=400, mesh=[9, 7, 11], box=[3.0, 3.2, 3.4],
         kind="DefaultNoChi", sigma=0.4, kappa=0.03, chi=[]),
    dict(name="sq_cubic", names=["A"], n=300, mesh=[8, 8, 8], box=[3.0, 3.0, 3.0],
         kind="SquaredPhi", sigma=0.6, kappa=0.1, chi=[]),
    dict(name="pme4", names=["A", "B", "C", "W"], n=800, mesh=[10, 12, 8], box=[3.5, 4.0, 3.0],
         kind="DefaultWithChi", sigma=0.5, kappa=0.05,
         chi=[("A", "B", 20.0), ("A", "C", -5.0), ("B", "C", 10.0), ("A", "W", 30.0), ("B", "W", 5.0)],
         coulombtype="PIC_Spectral", dielectric_const=80.0, type_charges=[1.0, -1.0, 0.0, 0.0]),
]


def setup(case):
    cfg = make_config(case["names"], case["n"], case["mesh"], case["box"], sigma=case["sigma"],
                      kappa=case["kappa"], hamiltonian=case["kind"], chi=case["chi"],
                      coulombtype=case.get("coulombtype"), dielectric_const=case.get("dielectric_const"))
    cfg.box_size = np.asarray(cfg.box_size, dtype=np.float64)
    cfg.pressure = True
    if case.get("type_charges") is not None:
        cfg.type_charges = list(case["type_charges"])
    return cfg


def close(x, ref, tol=1e-11):
    ref = np.asarray(ref)
    np.testing.assert_allclose(np.asarray(x), ref, rtol=0, atol=tol * max(np.abs(ref).max(), 1e-300))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_the_reference_field_module(case):
    pre = "field/" + case["name"]
    cfg = setup(case)
    pos, types, vel = G[pre + "/pos"], G[pre + "/types"], G[pre + "/vel"]
    T = cfg.n_types
    if case.get("coulombtype"):
        cfg.self_energy = fo.compute_self_energy_q(cfg, G[pre + "/charges"])
        assert cfg.self_energy == pytest.approx(float(G[pre + "/self_energy"]), rel=1e-14)
    W = OracleHamiltonian(cfg)
    st = fo.FieldState(cfg, np.float64)
    fo.update_field(st, W, pos, types, cfg, compute_potential=True)
    force = fo.compute_field_force(st, pos, types, T)
    close(force, G[pre + "/force"])
    close(np.stack(st.phi), G[pre + "/phi"])
    close(np.stack(st.v_ext), G[pre + "/v_ext"])
    close(np.stack(st.phi_fourier), G[pre + "/phi_fourier"])
    close(np.stack([np.stack(row) for row in st.force_mesh]), G[pre + "/force_mesh"])
    if case.get("coulombtype"):
        fq = fo.update_field_force_q(st, W, G[pre + "/charges"], pos, cfg)
        close(fq, G[pre + "/elec_forces"])
        close(st.psi, G[pre + "/psi"])
        close(st.phi_q, G[pre + "/phi_q"])
    e = fo.compute_field_and_kinetic_energy(st, W, vel, cfg)
    close(np.array(e), G[pre + "/energies"])
    p = fo.comp_pressure(st, W, vel, cfg, bond_pr=[1.0, -2.0, 0.5], angle_pr=[0.25, 0.5, -1.0])
    close(np.stack([np.stack(row) for row in st.phi_laplacian]), G[pre + "/phi_laplacian"])
    close(p, G[pre + "/pressure"], tol=1e-10)

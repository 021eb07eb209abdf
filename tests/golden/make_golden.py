"""Generate golden vectors from the REFERENCE's own importable modules.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

``hymd/hamiltonian.py`` imports only numpy + sympy, so it is loaded by file path
(the ``hymd`` package itself cannot be imported: h5py/mpi4py/pmesh are absent).
Outputs ``tests/golden/hamiltonian_golden.npz``: for each of the three
functionals and several parameter sets, the reference's ``H``, ``v_ext[t]``,
``w_0``, ``w_elec`` and the pressure helpers ``V_bar_0[t]`` / ``V_bar[t]`` evaluated on seeded
random inputs.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/hymd/hamiltonian.py"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from hymd_b200.config import Chi, Config  # noqa: E402


def load_reference_hamiltonian():
    spec = importlib.util.spec_from_file_location("ref_hamiltonian", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CASES = [
    dict(name="nochi_2", kind="DefaultNoChi", names=["A", "B"], chi=[], kappa=0.029230985982,
         sigma=0.2988365823859701, box=[7.1598, 11.2498, 5.1009], n=5),
    dict(name="nochi_2_f32params", kind="DefaultNoChi", names=["A", "B"], chi=[],
         kappa=0.029230985982, sigma=0.2988365823859701, box=[7.1598, 11.2498, 5.1009], n=5,
         f32_params=True),
    dict(name="sq_1", kind="SquaredPhi", names=["A"], chi=[], kappa=1.299759825895,
         sigma=1.2095870248085025, box=[15.0, 15.0, 15.0], n=3),
    dict(name="chi_3", kind="DefaultWithChi", names=["A", "B", "C"],
         chi=[("A", "B", 9.6754032616815161), ("A", "C", -13.2596290315913623),
              ("B", "C", 0.3852001771213374)], kappa=0.05, sigma=0.5,
         box=[15.0, 15.0, 15.0], n=5),
    dict(name="chi_5_dppc", kind="DefaultWithChi", names=["N", "P", "G", "C", "W"],
         chi=[("C", "W", 42.24), ("G", "C", 10.47), ("N", "W", -3.77), ("G", "W", 4.53),
              ("N", "P", -9.34), ("P", "G", 8.04), ("N", "G", 1.97), ("P", "C", 14.72),
              ("P", "W", -1.51), ("N", "C", 13.56)], kappa=0.05, sigma=0.5,
         box=[13.0, 13.0, 14.0], n=20000),
    dict(name="chi_4_pme", kind="DefaultWithChi", names=["A", "B", "C", "W"],
         chi=[("A", "B", 20.0), ("A", "C", -5.0), ("B", "C", 10.0), ("A", "W", 30.0),
              ("B", "W", 5.0)], kappa=0.05, sigma=0.5, box=[10.0, 10.0, 10.0], n=8370,
         coulombtype="PIC_Spectral", dielectric_const=80.0, self_energy=123.456,
         type_charges=[1.0, -1.0, 0.5, 0.0]),
]


def main():
    ref = load_reference_hamiltonian()
    out = {}
    meta = []
    rng = np.random.default_rng(20261017)
    for case in CASES:
        cfg = Config(mesh_size=16, sigma=case["sigma"], kappa=case["kappa"],
                     box_size=case["box"], hamiltonian=case["kind"],
                     chi=[Chi(*c) for c in case["chi"]],
                     coulombtype=case.get("coulombtype"),
                     dielectric_const=case.get("dielectric_const"),
                     self_energy=case.get("self_energy"))
        cfg.finalize(case["names"], n_particles=case["n"])
        # the reference reads plain attributes; give it a namespace copy so that it can
        # write rho0/a/simulation_volume the way Hamiltonian._setup does
        ns = types.SimpleNamespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
        ns.coulomb_constant = Config.coulomb_constant
        if case.get("type_charges") is not None:
            ns.type_charges = list(case["type_charges"])
        if not case.get("f32_params"):
            # float64 copy of the (float32-representable) box: pins the algebra exactly.
            # With the float32 box the reference derives a float32 rho0 and sympy then prints
            # `a` with ~7 significant digits into the generated lambda (numpy-2 behaviour of
            # hamiltonian.py:41-47); the "f32_params" case records that quirk.
            ns.box_size = np.asarray(cfg.box_size, dtype=np.float64)
        h = ref.get_hamiltonian(ns)
        t = cfg.n_types
        phi = [rng.uniform(0.0, 2.0 * ns.rho0, size=(4, 3, 5)) for _ in range(t)]
        k = [rng.normal(size=(4, 1, 1)) * 3, rng.normal(size=(1, 3, 1)) * 3,
             rng.normal(size=(1, 1, 5)) * 3]
        v = rng.normal(size=(4, 3, 5)) + 1j * rng.normal(size=(4, 3, 5))
        phi_q = rng.normal(size=(4, 3, 5))
        psi = rng.normal(size=(4, 3, 5))
        pre = case["name"]
        out[pre + "/phi"] = np.stack(phi)
        out[pre + "/k0"], out[pre + "/k1"], out[pre + "/k2"] = k
        out[pre + "/v"] = v
        out[pre + "/phi_q"], out[pre + "/psi"] = phi_q, psi
        out[pre + "/H"] = h.H(k, v)
        out[pre + "/w_0"] = np.asarray(h.w_0(phi), dtype=np.float64) * np.ones((4, 3, 5))
        out[pre + "/v_ext"] = np.stack(
            [np.asarray(h.v_ext[i](phi), dtype=np.float64) * np.ones((4, 3, 5)) for i in range(t)])
        out[pre + "/w_elec"] = np.asarray(h.w_elec([phi_q, psi]), dtype=np.float64)
        out[pre + "/V_bar_0"] = np.stack(
            [np.asarray(h.V_bar_0[i](phi), dtype=np.float64) * np.ones((4, 3, 5)) for i in range(t)])
        out[pre + "/V_bar"] = np.stack(
            [np.asarray(h.V_bar[i]([phi, psi]), dtype=np.float64) * np.ones((4, 3, 5)) for i in range(t)])
        out[pre + "/rho0_a_vol"] = np.array([ns.rho0, ns.a, ns.simulation_volume])
        meta.append({k2: v2 for k2, v2 in case.items()})
    np.savez_compressed(os.path.join(HERE, "hamiltonian_golden.npz"), **out)
    with open(os.path.join(HERE, "hamiltonian_golden.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print("wrote", len(out), "arrays for", len(CASES), "cases")


if __name__ == "__main__":
    main()

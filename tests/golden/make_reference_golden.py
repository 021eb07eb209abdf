"""Golden vectors produced by EXECUTING the reference's own Python source in the build container.

    python tests/golden/make_reference_golden.py        (needs /root/reference)

``tests/golden/ref_loader.py`` imports the unmodified ``hymd/force.py``, ``hymd/thermostat.py``,
``hymd/field.py``, ``hymd/pressure.py``, ``hymd/hamiltonian.py`` and ``hymd/input_parser.py`` from
``/root/reference`` with inert stand-ins for the missing third-party modules (one-rank MPI, and
``oracle/pmesh_standin.py`` as ``pmesh.pm``).  Outputs (committed):

* ``bonded_golden.npz``   -- the fixtures of the reference's ``test/conftest.py`` (``dppc_single``,
  ``alanine_octapeptide``), the term lists ``prepare_bonds`` builds for them, and the results of
  ``compute_{bond,angle,dihedral}_forces__plain`` per term (what ``test/test_force.py`` asserts) and
  for whole lists; plus seeded random chain systems in a periodic box.
* ``thermostat_golden.npz`` -- ``csvr_thermostat`` / ``cancel_com_momentum`` on the
  ``molecules_with_solvent`` fixture (``test/test_thermostat.py``) and on seeded random velocities,
  with prescribed random draws.
* ``field_golden.npz``    -- ``update_field`` + ``compute_field_force`` + ``update_field_force_q`` +
  ``compute_field_and_kinetic_energy`` + ``comp_pressure`` of the real ``hymd/field.py`` /
  ``hymd/pressure.py`` running on the pmesh stand-in, for small seeded systems.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import ref_loader as rl  # noqa: E402


def chain_system(rng, n_chains, chain_len, box, compact=False):
    """Random-walk chains of bond length ~0.47 nm; wrapped into the box unless ``compact``."""
    box = np.asarray(box, dtype=np.float64)
    r = np.empty((n_chains * chain_len, 3))
    for m in range(n_chains):
        start = (0.35 + 0.3 * rng.random(3)) * box if compact else rng.random(3) * box
        steps = rng.normal(size=(chain_len - 1, 3))
        steps *= (0.47 + 0.05 * rng.normal(size=(chain_len - 1, 1))) / np.linalg.norm(steps, axis=1)[:, None]
        if compact:
            steps *= 0.5
        r[m * chain_len:(m + 1) * chain_len] = start + np.concatenate([np.zeros((1, 3)), np.cumsum(steps, 0)])
    if not compact:
        r = np.mod(r, box)
    first = (np.arange(n_chains) * chain_len)[:, None]
    a2 = (first + np.arange(chain_len - 1)[None, :]).ravel()
    a3 = (first + np.arange(chain_len - 2)[None, :]).ravel()
    a4 = (first + np.arange(chain_len - 3)[None, :]).ravel()
    return r, a2, a3, a4


def bonded(out):
    f = rl.ref("force")
    ip = rl.ref("input_parser")
    # --- the reference's own fixtures -------------------------------------------------------
    indices, bonds, names, molecules, r, CONF = rl.conftest_fixture("dppc_single")
    config = ip.Config(n_steps=1, time_step=0.03, mesh_size=[30, 30, 30],
                       box_size=np.array([13.0, 13.0, 14.0]), sigma=0.5, kappa=1)
    config.bonds = CONF["bond_2"]
    config.angle_bonds = CONF["bond_3"]
    prep = f.prepare_bonds(molecules, names, bonds, indices, config)
    b2 = [np.asarray(x) for x in prep[0:4]]
    b3 = [np.asarray(x) for x in prep[4:9]]
    out["dppc/r"] = r
    out["dppc/box"] = np.array(CONF["L"], dtype=np.float64)
    for key, arr in zip(("a", "b", "r0", "k"), b2):
        out["dppc/b2_" + key] = arr
    for key, arr in zip(("a", "b", "c", "t0", "k"), b3):
        out["dppc/b3_" + key] = arr
    e_terms, f_terms = [], []
    for i in range(len(b2[0])):
        fb = np.zeros_like(r)
        e_terms.append(f.compute_bond_forces__plain(fb, r.copy(), [[x[i] for x in b2]], CONF["L"]))
        f_terms.append(fb)
    out["dppc/b2_term_energy"], out["dppc/b2_term_force"] = np.array(e_terms), np.array(f_terms)
    fb = np.zeros_like(r)
    out["dppc/b2_energy"] = f.compute_bond_forces__plain(fb, r.copy(), list(zip(*b2)), CONF["L"])
    out["dppc/b2_force"] = fb
    e_terms, f_terms = [], []
    for i in range(len(b3[0])):
        fa = np.zeros_like(r)
        e_terms.append(f.compute_angle_forces__plain(fa, r.copy(), [[x[i] for x in b3]], CONF["L"]))
        f_terms.append(fa)
    out["dppc/b3_term_energy"], out["dppc/b3_term_force"] = np.array(e_terms), np.array(f_terms)
    fa = np.zeros_like(r)
    out["dppc/b3_energy"] = f.compute_angle_forces__plain(fa, r.copy(), list(zip(*b3)), CONF["L"])
    out["dppc/b3_force"] = fa

    indices, bonds, names, molecules, r, CONF = rl.conftest_fixture("alanine_octapeptide")
    config = ip.Config(n_steps=1, time_step=0.03, mesh_size=[30, 30, 30],
                       box_size=np.array([5.0, 5.0, 5.0]), sigma=0.5, kappa=1)
    config.dihedrals = CONF["bond_4"]
    prep = f.prepare_bonds(molecules, names, bonds, indices, config)
    a, b, c, d, coeff, dtype, last = (np.asarray(x) for x in prep[9:16])
    out["ala/r"], out["ala/box"] = r, np.array(CONF["L"], dtype=np.float64)
    out["ala/a"], out["ala/b"], out["ala/c"], out["ala/d"] = a, b, c, d
    out["ala/coeff"], out["ala/dtype"], out["ala/last"] = coeff, dtype, last
    e_terms, f_terms = [], []
    for i in range(len(a)):
        fd = np.zeros_like(r)
        e_terms.append(f.compute_dihedral_forces__plain(
            fd, r.copy(), [[a[i], b[i], c[i], d[i], coeff[i, 0:2], 0]], CONF["L"]))
        f_terms.append(fd)
    out["ala/term_energy"], out["ala/term_force_plain"] = np.array(e_terms), np.array(f_terms)

    # --- seeded random chains ---------------------------------------------------------------
    rng = np.random.default_rng(4101)
    box = np.array([6.0, 7.5, 5.5])
    r, a2, a3, a4 = chain_system(rng, 40, 12, box)
    r0 = 0.47 + 0.1 * rng.random(len(a2))
    k2 = 1250.0 * (0.5 + rng.random(len(a2)))
    t0 = np.radians(rng.choice([120.0, 180.0, 140.0], size=len(a3)))
    k3 = 25.0 * (0.5 + rng.random(len(a3)))
    out["chains/r"], out["chains/box"] = r, box
    out["chains/b2_a"], out["chains/b2_b"], out["chains/b2_r0"], out["chains/b2_k"] = a2, a2 + 1, r0, k2
    out["chains/b3_a"], out["chains/b3_b"], out["chains/b3_c"] = a3, a3 + 1, a3 + 2
    out["chains/b3_t0"], out["chains/b3_k"] = t0, k3
    fb = np.zeros_like(r)
    out["chains/b2_energy"] = f.compute_bond_forces__plain(
        fb, r.copy(), list(zip(a2, a2 + 1, r0, k2)), box)
    out["chains/b2_force"] = fb
    fa = np.zeros_like(r)
    out["chains/b3_energy"] = f.compute_angle_forces__plain(
        fa, r.copy(), list(zip(a3, a3 + 1, a3 + 2, t0, k3)), box)
    out["chains/b3_force"] = fa
    # dihedrals: compact chains that never cross a box face (the deprecated plain function
    # subtracts the per-dimension image shift from all three components, force.py:822-825)
    r, _, _, a4 = chain_system(rng, 25, 10, box, compact=True)
    coeff = np.zeros((len(a4), 6, 5))
    coeff[:, 0] = rng.normal(size=(len(a4), 5)) * 4.0
    coeff[:, 1] = rng.uniform(-np.pi, np.pi, size=(len(a4), 5))
    out["dih/r"], out["dih/box"] = r, box
    out["dih/a"], out["dih/b"], out["dih/c"], out["dih/d"] = a4, a4 + 1, a4 + 2, a4 + 3
    out["dih/coeff"] = coeff
    fd = np.zeros_like(r)
    out["dih/energy"] = f.compute_dihedral_forces__plain(
        fd, r.copy(), [[a4[i], a4[i] + 1, a4[i] + 2, a4[i] + 3, coeff[i, 0:2], 0] for i in range(len(a4))], box)
    out["dih/force_plain"] = fd


class _Mock:
    def __init__(self, x):
        self.x, self.i = list(x), 0

    def __call__(self, *args):
        self.i += 1
        return self.x[self.i - 1]


def thermostat(out):
    th = rl.ref("thermostat")
    ip = rl.ref("input_parser")
    comm = sys.modules["mpi4py"].MPI.COMM_WORLD
    (indices, positions, molecules, velocities, bonds, names, types_) = \
        rl.conftest_fixture("molecules_with_solvent")
    out["mws/velocities"] = velocities
    out["mws/names"] = names
    base = dict(n_steps=0, time_step=0.032958582578275, box_size=np.array([10.0, 10.0, 10.0]),
                tau=0.925852989520023, mesh_size=[2, 2, 2], sigma=0.5, kappa=0.05,
                n_particles=len(indices), target_temperature=310.0, thermostat_work=0.0, mass=72.0)
    cases = [
        ("all_nocom", [], False, [0.5579657512081987], [125.4595634810623]),
        ("abcd_nocom", [["A"], ["B"], ["C"], ["D"]], False,
         [-1.752320325907187, 1.099694957420625, 0.6113448515745533, -0.7183266831611322],
         [35.43824087713971, 27.57022975113815, 2.024725328228174, 36.50208472031436]),
        ("abc_d_nocom", [["A", "B", "C"], ["D"]], False,
         [0.1661408606772054, -0.06216797747541603], [97.07130218590895, 33.06359496718739]),
        ("all_com", [], True, [0.5579657512081987], [125.4595634810623]),
        ("abcd_com", [["A"], ["B"], ["C"], ["D"]], True,
         [-1.752320325907187, 1.099694957420625, 0.6113448515745533, -0.7183266831611322],
         [35.43824087713971, 27.57022975113815, 2.024725328228174, 36.50208472031436]),
        ("abc_d_com", [["A", "B", "C"], ["D"]], True,
         [0.1661408606772054, -0.06216797747541603], [97.07130218590895, 33.06359496718739]),
    ]
    for name, groups, remove, gauss, chi2 in cases:
        config = ip.Config(**base)
        config = ip._find_unique_names(config, names, comm=comm)
        config.thermostat_coupling_groups = [list(g) for g in groups]
        v = velocities.copy()
        th.csvr_thermostat(v, names, config, np.random.default_rng(), comm=comm,
                           random_gaussian=_Mock(gauss), random_chi_squared=_Mock(chi2),
                           remove_center_of_mass_momentum=remove)
        out[f"mws/{name}/v"] = v
        out[f"mws/{name}/work"] = np.float64(config.thermostat_work)
        out[f"mws/{name}/gauss"], out[f"mws/{name}/chi2"] = np.array(gauss), np.array(chi2)
    config = ip.Config(**base)
    out["mws/cancel_com"] = th.cancel_com_momentum(velocities.copy(), config, comm=comm)

    # seeded larger system, three groups of uneven size, default com removal
    rng = np.random.default_rng(4102)
    n = 5000
    names_l = rng.choice(np.array([b"A", b"B", b"W"], dtype="S5"), size=n, p=[0.2, 0.3, 0.5])
    v0 = rng.normal(scale=0.2, size=(n, 3)) + np.array([0.01, -0.02, 0.005])
    gauss = rng.normal(size=2)
    chi2 = np.array([rng.chisquare(3 * np.sum(names_l != b"W") - 1),
                     rng.chisquare(3 * np.sum(names_l == b"W") - 1)])
    config = ip.Config(**{**base, "n_particles": n, "respa_inner": 5, "tau": 0.1})
    config = ip._find_unique_names(config, names_l, comm=comm)
    config.thermostat_coupling_groups = [["A", "B"], ["W"]]
    v = v0.copy()
    th.csvr_thermostat(v, names_l, config, np.random.default_rng(), comm=comm,
                       random_gaussian=_Mock(gauss), random_chi_squared=_Mock(chi2))
    out["rand/names"], out["rand/v0"], out["rand/v"] = names_l, v0, v
    out["rand/work"] = np.float64(config.thermostat_work)
    out["rand/gauss"], out["rand/chi2"] = gauss, chi2
    out["rand/params"] = np.array([config.mass, config.gas_constant, config.target_temperature,
                                   config.time_step, config.respa_inner, config.tau])
    out["mws/params"] = np.array([72.0, ip.Config.gas_constant, 310.0, base["time_step"], 1,
                                  base["tau"]])


FIELD_CASES = [
    dict(name="chi3_even", names=["A", "B", "C"], frac=[0.3, 0.3, 0.4], n=600, mesh=[12, 10, 8],
         box=[4.0, 3.5, 3.0], kind="DefaultWithChi", sigma=0.5, kappa=0.05,
         chi=[("A", "B", 20.0), ("A", "C", -5.0), ("B", "C", 10.0)]),
    dict(name="nochi_odd", names=["A", "B"], frac=[0.5, 0.5], n=400, mesh=[9, 7, 11],
         box=[3.0, 3.2, 3.4], kind="DefaultNoChi", sigma=0.4, kappa=0.03, chi=[]),
    dict(name="sq_cubic", names=["A"], frac=[1.0], n=300, mesh=[8, 8, 8],
         box=[3.0, 3.0, 3.0], kind="SquaredPhi", sigma=0.6, kappa=0.1, chi=[]),
    dict(name="pme4", names=["A", "B", "C", "W"], frac=[0.2, 0.2, 0.1, 0.5], n=800, mesh=[10, 12, 8],
         box=[3.5, 4.0, 3.0], kind="DefaultWithChi", sigma=0.5, kappa=0.05,
         chi=[("A", "B", 20.0), ("A", "C", -5.0), ("B", "C", 10.0), ("A", "W", 30.0),
              ("B", "W", 5.0)], coulombtype="PIC_Spectral", dielectric_const=80.0,
         type_charges=[1.0, -1.0, 0.0, 0.0]),
]


def field(out):
    fd = rl.ref("field")
    pr = rl.ref("pressure")
    hm = rl.ref("hamiltonian")
    pmesh = sys.modules["pmesh.pm"]
    comm = sys.modules["mpi4py"].MPI.COMM_WORLD
    from hymd_b200.config import Chi, Config
    for ci, case in enumerate(FIELD_CASES):
        rng = np.random.default_rng(4200 + ci)
        n, T = case["n"], len(case["names"])
        cfg = Config(mesh_size=case["mesh"], sigma=case["sigma"], kappa=case["kappa"],
                     box_size=case["box"], hamiltonian=case["kind"],
                     chi=[Chi(*c) for c in case["chi"]], coulombtype=case.get("coulombtype"),
                     dielectric_const=case.get("dielectric_const"), dtype=np.float64)
        cfg.finalize(case["names"], n_particles=n)
        ns = types.SimpleNamespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
        ns.coulomb_constant, ns.gas_constant = Config.coulomb_constant, Config.gas_constant
        ns.box_size = np.asarray(cfg.box_size, dtype=np.float64)
        ns.pressure = True
        if case.get("type_charges") is not None:
            ns.type_charges = np.array(case["type_charges"])
        types_ = rng.choice(T, size=n, p=case["frac"]).astype(np.int64)
        pos = rng.random((n, 3)) * ns.box_size
        vel = rng.normal(scale=0.2, size=(n, 3))
        charges = None
        if case.get("coulombtype"):
            charges = np.asarray(ns.type_charges)[types_].astype(np.float64)
            ns.self_energy = fd.compute_self_energy_q(ns, charges, comm=comm)
        h = hm.get_hamiltonian(ns)
        pm, field_list, elec_common, coulomb = fd.initialize_pm(pmesh, ns, comm=comm)
        phi, phi_fourier, force_on_grid, v_ext_fourier, v_ext, phi_transfer, phi_laplacian = field_list
        layouts = [pm.decompose(pos[types_ == t]) for t in range(T)]
        fd.update_field(phi, phi_laplacian, phi_transfer, layouts, force_on_grid, h, pm, pos, types_,
                        ns, v_ext, phi_fourier, v_ext_fourier, ns.m, compute_potential=True)
        force = np.zeros((n, 3))
        fd.compute_field_force(layouts, pos, force_on_grid, force, types_, T)
        phi_q = psi = None
        pre = "field/" + case["name"]
        if charges is not None:
            phi_q, phi_q_fourier, psi, elec_field = elec_common
            elec_field_fourier, psi_fourier = coulomb
            elec_forces = np.zeros((n, 3))
            fd.update_field_force_q(charges, phi_q, phi_q_fourier, psi, psi_fourier, elec_field_fourier,
                                    elec_field, elec_forces, pm.decompose(pos), h, pm, pos, ns)
            out[pre + "/elec_forces"], out[pre + "/psi"] = elec_forces, np.asarray(psi)
            out[pre + "/phi_q"], out[pre + "/charges"] = np.asarray(phi_q), charges
            out[pre + "/self_energy"] = np.float64(ns.self_energy)
        e = fd.compute_field_and_kinetic_energy(phi, phi_q, psi, vel, h, pos, types_, v_ext, ns,
                                                layouts, comm=comm)
        p = pr.comp_pressure(phi, phi_q, psi, h, vel, ns, phi_fourier, phi_laplacian, phi_transfer,
                             pos, np.array([1.0, -2.0, 0.5]), np.array([0.25, 0.5, -1.0]), comm=comm)
        out[pre + "/pos"], out[pre + "/types"], out[pre + "/vel"] = pos, types_, vel
        out[pre + "/force"] = force
        out[pre + "/phi"] = np.stack([np.asarray(x) for x in phi])
        out[pre + "/v_ext"] = np.stack([np.asarray(x) for x in v_ext])
        out[pre + "/phi_fourier"] = np.stack([np.asarray(x) for x in phi_fourier])
        out[pre + "/force_mesh"] = np.stack([np.stack([np.asarray(x) for x in row]) for row in force_on_grid])
        out[pre + "/phi_laplacian"] = np.stack([np.stack([np.asarray(x) for x in row]) for row in phi_laplacian])
        out[pre + "/energies"] = np.array([float(x) for x in e])
        out[pre + "/pressure"] = np.asarray(p, dtype=np.float64)


GPE_CASES = [
    dict(name="gpe3", names=["A", "B", "W"], frac=[0.25, 0.25, 0.5], n=900, mesh=[10, 12, 8],
         box=[3.5, 4.0, 3.0], sigma=0.5, kappa=0.05, chi=[("A", "B", 15.0), ("A", "W", 25.0)],
         type_charges=[1.0, -1.0, 0.0], dielectric_type=[5.0, 10.0, 80.0], pol_mixing=0.6, conv_crit=1e-6),
    dict(name="gpe2_odd", names=["A", "W"], frac=[0.4, 0.6], n=700, mesh=[9, 8, 11],
         box=[3.0, 3.2, 3.4], sigma=0.45, kappa=0.05, chi=[("A", "W", 10.0)],
         type_charges=[0.5, -0.3333333333333333], dielectric_type=[20.0, 60.0], pol_mixing=0.5,
         conv_crit=1e-7),
]


def gpe(out):
    """update_field_force_q_GPE / compute_field_energy_q_GPE (field.py:964-1112, 706-760) of the real
    hymd/field.py over the pmesh stand-in: groundwork for SURVEY.md section 8 row f3."""
    fd = rl.ref("field")
    hm = rl.ref("hamiltonian")
    pmesh = sys.modules["pmesh.pm"]
    comm = sys.modules["mpi4py"].MPI.COMM_WORLD
    from hymd_b200.config import Chi, Config
    for ci, case in enumerate(GPE_CASES):
        rng = np.random.default_rng(4300 + ci)
        n, T = case["n"], len(case["names"])
        cfg = Config(mesh_size=case["mesh"], sigma=case["sigma"], kappa=case["kappa"], box_size=case["box"],
                     hamiltonian="DefaultWithChi", chi=[Chi(*c) for c in case["chi"]],
                     coulombtype="PIC_Spectral_GPE", dtype=np.float64)
        cfg.finalize(case["names"], n_particles=n)
        ns = types.SimpleNamespace(**{k: getattr(cfg, k) for k in cfg.__dataclass_fields__})
        ns.coulomb_constant, ns.gas_constant = Config.coulomb_constant, Config.gas_constant
        ns.box_size = np.asarray(cfg.box_size, dtype=np.float64)
        ns.type_charges = np.array(case["type_charges"])
        ns.dielectric_type = np.array(case["dielectric_type"])
        ns.pol_mixing, ns.conv_crit = case["pol_mixing"], case["conv_crit"]
        types_ = rng.choice(T, size=n, p=case["frac"]).astype(np.int64)
        pos = rng.random((n, 3)) * ns.box_size
        charges = ns.type_charges[types_].astype(np.float64)
        h = hm.get_hamiltonian(ns)
        pm, field_list, elec_common, coulomb = fd.initialize_pm(pmesh, ns, comm=comm)
        phi, phi_fourier, force_on_grid, v_ext_fourier, v_ext, phi_transfer, phi_laplacian = field_list
        phi_q, phi_q_fourier, psi, elec_field = elec_common
        (phi_eps, phi_eps_fourier, phi_eta, phi_eta_fourier, phi_pol, phi_pol_prev, elec_dot,
         elec_field_contrib, Vbar_elec, Vbar_elec_fourier, force_mesh_elec, force_mesh_elec_fourier) = coulomb
        layouts = [pm.decompose(pos[types_ == t]) for t in range(T)]
        fd.update_field(phi, phi_laplacian, phi_transfer, layouts, force_on_grid, h, pm, pos, types_, ns,
                        v_ext, phi_fourier, v_ext_fourier, ns.m)

        def conv_fun(comm_, diffmesh):        # main.py:158-163 (default "max_diff")
            return comm_.allreduce(np.max(diffmesh), op="MAX")
        elec_forces = np.zeros((n, 3))
        Vbar, eps, dot = fd.update_field_force_q_GPE(
            conv_fun, phi, types_, charges, phi_q, phi_q_fourier, phi_eps, phi_eps_fourier, phi_eta,
            phi_eta_fourier, phi_pol_prev, phi_pol, elec_field, elec_forces, elec_field_contrib, psi, Vbar_elec,
            Vbar_elec_fourier, force_mesh_elec, force_mesh_elec_fourier, h, pm.decompose(pos), layouts, pm,
            pos, ns, comm=comm)
        energy = fd.compute_field_energy_q_GPE(ns, eps, 0.0, dot, comm=comm)
        pre = "gpe/" + case["name"]
        out[pre + "/pos"], out[pre + "/types"], out[pre + "/charges"] = pos, types_, charges
        out[pre + "/phi"] = np.stack([np.asarray(x) for x in phi])
        out[pre + "/elec_forces"] = elec_forces
        out[pre + "/psi"], out[pre + "/phi_eps"] = np.asarray(psi), np.asarray(eps)
        out[pre + "/elec_dot"] = np.asarray(dot)
        out[pre + "/Vbar_elec"] = np.stack([np.asarray(x) for x in Vbar])
        out[pre + "/energy"] = np.float64(energy)


BAROSTAT_PRESSURE = np.array([3.1, -2.0, 0.4, 0.1, 0.2, 0.3, 0.01, 0.02, 0.03, -0.01, 0.0, 0.02, 0.0, 0.0, 0.0,
                              1.7, 2.3, -0.9])        # what the patched comp_pressure returns


def barostat(out):
    """isotropic / semiisotropic of hymd/barostat.py (Berendsen) and hymd/barostat_scr.py (SCR) with
    ``comp_pressure`` patched to return BAROSTAT_PRESSURE and ``initialize_pm`` patched to a marker: the
    scaling arithmetic, the prng call order and the in-place updates of box and positions are the
    reference's own."""
    ip = rl.ref("input_parser")
    mods = {"berendsen": rl.ref("barostat"), "scr": rl.ref("barostat_scr")}
    comm = sys.modules["mpi4py"].MPI.COMM_WORLD
    rng = np.random.default_rng(4400)
    pos0 = rng.random((7, 3)) * np.array([5.0, 6.0, 7.0])
    out["barostat/pressure"], out["barostat/pos0"] = BAROSTAT_PRESSURE, pos0
    for kind, mod in mods.items():
        mod.comp_pressure = lambda *a, **k: BAROSTAT_PRESSURE.copy()
        mod.initialize_pm = lambda pmesh, config, comm=None: "reinitialized"
        for fn in ("isotropic", "semiisotropic"):
            for step, (P_L, P_N) in ((4, (1.0, 1.0)), (4, (1.0, None)), (5, (1.0, 1.0))):
                config = ip.Config(n_steps=1, time_step=0.03, mesh_size=[8, 8, 8], sigma=0.5, kappa=0.05,
                                   box_size=np.array([5.0, 6.0, 7.0]), respa_inner=5, n_b=2, tau_p=1.5,
                                   target_temperature=323.0)
                config.target_pressure = mod.Target_pressure(P_L=P_L, P_N=P_N)
                pos = pos0.copy()
                res, change = getattr(mod, fn)(None, "old", None, None, None, None, pos, None, config, None,
                                               None, None, np.zeros(3), np.zeros(3), step,
                                               np.random.default_rng(99), comm=comm)
                pre = f"barostat/{kind}/{fn}/{step}_{P_L}_{P_N}"
                out[pre + "/box"], out[pre + "/pos"] = np.asarray(config.box_size, dtype=np.float64), pos
                out[pre + "/change"] = np.array([bool(change), res == "reinitialized"])
                out[pre + "/surface_tension"] = np.float64(getattr(config, "surface_tension", None) or 0.0)


def main():
    for fn, name in ((bonded, "bonded_golden.npz"), (thermostat, "thermostat_golden.npz"),
                     (field, "field_golden.npz"), (gpe, "gpe_golden.npz"), (barostat, "barostat_golden.npz")):
        out = {}
        fn(out)
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, len(out), "arrays", os.path.getsize(os.path.join(HERE, name)) // 1024, "KiB")


if __name__ == "__main__":
    main()

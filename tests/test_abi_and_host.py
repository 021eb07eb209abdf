"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares,
the host-side mirror of the reference interface behaves, synthetic systems are well formed."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

from conftest import ROOT, make_config


def _header_functions():
    src = open(os.path.join(ROOT, "include", "hymd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hymd_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hymd_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hymd_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert _lib.load().hymd_abi_version() == 1


def _header_prototypes():
    """{name: (return type, [parameter declarations])} of every prototype in the header."""
    src = open(os.path.join(ROOT, "include", "hymd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for ret, name, params in re.findall(r"^\s*((?:const\s+)?[a-z_0-9]+\s*\*?)\s*(hymd_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", src,
                                        flags=re.M | re.S):
        plist = [q.strip() for q in params.split(",")] if params.strip() not in ("", "void") else []
        out[name] = (ret.strip(), plist)
    return out


def test_ctypes_binding_matches_the_header_prototypes():
    """Every prototype of include/hymd_b200.h against the argtypes / restype hymd_b200/_lib.py declares: same
    number of parameters, pointers bound as pointers, 64-bit integers as 64-bit, doubles as doubles (a binding
    that drifts from the header corrupts the call silently: ctypes checks nothing)."""
    from hymd_b200 import _lib
    lib = _lib.load()
    protos = _header_prototypes()
    assert sorted(protos) == sorted(_lib.EXPORTS)
    pointer_like = (ctypes.c_void_p, ctypes.c_char_p)
    for name, (ret, params) in protos.items():
        fn = getattr(lib, name)
        argtypes = fn.argtypes
        if argtypes is None:            # entry points the Python layer never calls with arguments
            assert not params or name in ("hymd_abi_version",), f"{name}: no argtypes declared"
            continue
        assert len(argtypes) == len(params), f"{name}: header has {len(params)} parameters, binding {len(argtypes)}"
        for decl, at in zip(params, argtypes):
            is_ptr = "*" in decl or "[" in decl
            bound_ptr = issubclass(at, pointer_like) or hasattr(at, "contents") or issubclass(at, ctypes.Array)
            assert is_ptr == bound_ptr, f"{name}: '{decl}' bound as {at.__name__}"
            if not is_ptr:
                base = decl.replace("const", "").split()[0]
                want = {"int": 4, "int32_t": 4, "int64_t": 8, "double": 8}[base]
                assert ctypes.sizeof(at) == want, f"{name}: '{decl}' bound as {at.__name__}"
                assert (base == "double") == (at is ctypes.c_double), f"{name}: '{decl}' bound as {at.__name__}"
        if ret == "int64_t":
            assert fn.restype is ctypes.c_int64, name
        elif ret.replace(" ", "") == "constchar*":
            assert fn.restype is ctypes.c_char_p, name


def test_config_struct_matches_header_layout():
    from hymd_b200 import _lib
    # 5 int32 (+4 pad) + 3 double + 4 int32 + 2 double + (32*32 + 32 + 32) double
    assert ctypes.sizeof(_lib.HymdConfig) == 24 + 24 + 16 + 16 + 8 * (1024 + 64)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hymd_b200 import _lib
    from hymd_b200.field import initialize_pm
    cfg = make_config(["A", "B"], 10, 8, [2.0, 2.0, 2.0])
    with pytest.raises(_lib.HymdError):
        initialize_pm(None, cfg)


def test_signatures_mirror_reference():
    """Positional argument names of the reference functions (field.py:10, 152, 203, 241-255,
    428-444, 619-631, 1115-1117)."""
    from hymd_b200 import field
    expect = {
        "initialize_pm": ["pmesh", "config", "comm"],
        "compute_field_force": ["layouts", "r", "force_mesh", "force", "types", "n_types"],
        "compute_self_energy_q": ["config", "charges", "comm"],
        "update_field_force_q": ["charges", "phi_q", "phi_q_fourier", "psi", "psi_fourier",
                                 "elec_field_fourier", "elec_field", "elec_forces", "layout_q",
                                 "hamiltonian", "pm", "positions", "config"],
        "update_field": ["phi", "phi_laplacian", "phi_transfer", "layouts", "force_mesh",
                         "hamiltonian", "pm", "positions", "types", "config", "v_ext",
                         "phi_fourier", "v_ext_fourier", "m", "compute_potential"],
        "compute_field_and_kinetic_energy": ["phi", "phi_q", "psi", "velocity", "hamiltonian",
                                             "positions", "types", "v_ext", "config", "layouts",
                                             "comm"],
    }
    for name, args in expect.items():
        got = list(inspect.signature(getattr(field, name)).parameters)
        assert got == args, name
    dd = inspect.signature(field.domain_decomposition)
    assert list(dd.parameters)[:3] == ["positions", "pm", "args"]
    for kw in ("molecules", "bonds", "topol", "verbose", "comm"):
        assert kw in dd.parameters


def test_affine_parameters_match_closed_form_and_reject_nonaffine():
    from hymd_b200.hamiltonian import affine_parameters, get_hamiltonian
    chi = [("A", "B", 9.6754032616815161), ("A", "C", -13.2596290315913623),
           ("B", "C", 0.3852001771213374)]
    cfg = make_config(["A", "B", "C"], 5, 16, [15.0, 15.0, 15.0], chi=chi)
    h = get_hamiltonian(cfg)
    A, c = affine_parameters(h, 3)
    k, r = cfg.kappa, cfg.rho0
    want = np.full((3, 3), 1.0 / (k * r)) + h.chi_matrix / r
    np.testing.assert_allclose(A, want, rtol=1e-13)
    np.testing.assert_allclose(c, -cfg.a / (k * r), rtol=1e-13)

    class Quartic:
        v_ext = [lambda phi: phi[0] ** 3]
    with pytest.raises(ValueError):
        affine_parameters(Quartic(), 1)


def test_hamiltonian_mirror_matches_oracle_and_golden():
    import json
    from hymd_b200.config import Chi, Config
    from hymd_b200.hamiltonian import get_hamiltonian
    gold = np.load(os.path.join(ROOT, "tests", "golden", "hamiltonian_golden.npz"))
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "hamiltonian_golden.json")))
    for case in cases:
        if case.get("f32_params"):
            continue
        cfg = Config(mesh_size=16, sigma=case["sigma"], kappa=case["kappa"], box_size=case["box"],
                     hamiltonian=case["kind"], chi=[Chi(*c) for c in case["chi"]],
                     coulombtype=case.get("coulombtype"),
                     dielectric_const=case.get("dielectric_const"),
                     self_energy=case.get("self_energy"))
        cfg.finalize(case["names"], n_particles=case["n"])
        cfg.box_size = np.asarray(cfg.box_size, dtype=np.float64)
        h = get_hamiltonian(cfg)
        pre = case["name"]
        phi = list(gold[pre + "/phi"])
        np.testing.assert_allclose(h.w_0(phi), gold[pre + "/w_0"], rtol=1e-11, atol=1e-9)
        for t in range(cfg.n_types):
            np.testing.assert_allclose(h.v_ext[t](phi), gold[pre + "/v_ext"][t], rtol=1e-11, atol=1e-9)
        k = [gold[pre + "/k0"], gold[pre + "/k1"], gold[pre + "/k2"]]
        np.testing.assert_allclose(h.H(k, gold[pre + "/v"]), gold[pre + "/H"], rtol=1e-13, atol=1e-300)


@pytest.mark.parametrize("name,n,mesh", [("C1", 2000, 12), ("C2", 5000, 16), ("C3", 4000, 16),
                                         ("C4", 3000, 16)])
def test_synthetic_systems(name, n, mesh):
    from hymd_b200.synthetic import make_system
    s = make_system(name, np.float32, n=n, mesh=mesh)
    L = float(s.config.box_size[0])
    assert s.positions.shape == (n, 3) and s.positions.dtype == np.float32
    assert (s.positions >= 0).all() and (s.positions < L).all()
    assert s.types.min() >= 0 and s.types.max() < s.config.n_types
    assert abs(n / L ** 3 - 8.37) < 0.05
    if s.charges is not None:
        assert abs(float(s.charges.sum())) < 1e-6
        assert s.config.coulombtype == "PIC_Spectral"
    s2 = make_system(name, np.float32, n=n, mesh=mesh)
    assert np.array_equal(s.positions, s2.positions)


def test_row_f2_signatures_mirror_reference():
    """Argument names of the f2py kernels (compute_bond_forces.f90:1, compute_angle_forces.f90:1,
    compute_dihedral_forces.f90:1), of thermostat.py:12, 18, 111-121, of barostat.py:43-61 and of
    pressure.py:84-98."""
    from hymd_b200 import barostat, force, pressure, thermostat
    expect = [
        (force.compute_bond_forces, ["f_bonds", "r", "box_size", "a", "b", "r0", "k"]),
        (force.compute_angle_forces, ["f_angles", "r", "box_size", "a", "b", "c", "t0", "k"]),
        (force.compute_dihedral_forces, ["f_dihedrals", "r", "dipoles", "transfer_matrix", "box_size", "a", "b",
                                         "c", "d", "coeff", "dtype", "bb_index", "dipole_flag"]),
        (thermostat.csvr_thermostat, ["velocity", "names", "config", "prng", "comm", "random_gaussian",
                                      "random_chi_squared", "remove_center_of_mass_momentum"]),
        (thermostat.cancel_com_momentum, ["velocities", "config", "comm"]),
        (thermostat.generate_initial_velocities, ["velocities", "config", "prng", "comm"]),
        (pressure.comp_pressure, ["phi", "phi_q", "psi", "hamiltonian", "velocities", "config", "phi_fourier",
                                  "phi_laplacian", "phi_transfer", "positions", "bond_pr", "angle_pr", "comm"]),
    ]
    baro = ["pmesh", "pm_stuff", "phi", "phi_q", "psi", "hamiltonian", "positions", "velocities", "config",
            "phi_fft", "phi_laplacian", "phi_transfer", "bond_pr", "angle_pr", "step", "prng", "comm"]
    for ns in (barostat.berendsen, barostat.scr):
        expect += [(ns.isotropic, baro), (ns.semiisotropic, baro)]
    for fn, args in expect:
        assert list(inspect.signature(fn).parameters) == args, fn.__name__


def test_product_never_imports_the_oracle_or_the_test_shim():
    """oracle/ and tests/native are test infrastructure: no module of the product may import them."""
    import re
    pkg = os.path.join(ROOT, "hymd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", text, re.M), f
                assert not re.search(r"#\s*include\s+[<\"][^>\"]*(oracle|tests)/", text), f
                assert "libhost_check" not in text, f


def test_row_f2_needs_a_gpu_and_says_so():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hymd_b200 import _lib, force, thermostat
    r = np.zeros((4, 3))
    a = np.arange(3)
    with pytest.raises(_lib.HymdError):
        force.compute_bond_forces(np.zeros_like(r), r, np.ones(3), a, a + 1, np.ones(3), np.ones(3))
    with pytest.raises(_lib.HymdError):
        thermostat.cancel_com_momentum(np.zeros((4, 3)), type("C", (), {"n_particles": 4})())


def test_block_ordering_keeps_molecules_whole_and_separates_chains_from_solvent():
    """HYMD_B200_DD_BLOCK (hymd_b200.field._cell_order with block > 0), on CPU tensors: a permutation,
    molecules contiguous with their internal order kept, and 128-particle tiles mostly homogeneous."""
    import types

    import numpy as np
    import torch
    from hymd_b200 import field as F
    pm = types.SimpleNamespace(Nmesh=np.array([32, 32, 32]), BoxSize=np.array([13.0, 13.0, 13.0]),
                               device=torch.device("cpu"))
    rng = np.random.default_rng(0)
    nch, length, nsol = 400, 10, 4000
    first = rng.random((nch, 1, 3)) * 13
    pos = torch.tensor(np.concatenate([(first + np.zeros((nch, length, 3))).reshape(-1, 3), rng.random((nsol, 3)) * 13]))
    mol = torch.tensor(np.concatenate([np.repeat(np.arange(nch), length), nch + np.arange(nsol)]))
    n = len(pos)
    for block in (0, 16):
        perm = F._cell_order(pm, pos, mol, block).numpy()
        assert sorted(perm.tolist()) == list(range(n))
        m = mol.numpy()[perm]
        starts = np.nonzero(np.diff(m, prepend=-1))[0]
        assert len(starts) == nch + nsol                       # every molecule is one contiguous run
        chain_rows = perm[m < nch].reshape(nch, length)
        assert (np.diff(chain_rows, axis=1) == 1).all()        # internal order kept
    single = (np.arange(n) >= nch * length)[perm]
    tiles = single[: n // 128 * 128].reshape(-1, 128).mean(axis=1)
    assert ((tiles > 0) & (tiles < 1)).mean() < 0.4            # cell order alone: every tile is mixed


def test_device_scalar_behaves_like_the_float_the_reference_expects():
    """config.thermostat_work is a float in the reference; file_io.py:751 subtracts it and
    file_io.py:704 stores it into an HDF5 dataset."""
    import torch
    from hymd_b200.thermostat import DeviceScalar
    w = DeviceScalar(torch.tensor([2.5], dtype=torch.float64))
    assert 10.0 - 1.0 - w == 6.5 and w - 1 == 1.5 and 2 * w == 5.0 and w / 2 == 1.25
    assert isinstance(10.0 - w, float) and -w == -2.5 and abs(w) == 2.5 and w < 3 and w >= 2.5
    a = np.zeros(2)
    a[1] = w
    assert a[1] == 2.5 and float(np.asarray(w)) == 2.5 and f"{w:.2f}" == "2.50"
    assert np.float64(1.0) + w == 3.5


def test_write_epoch_changes_the_position_fingerprint():
    import torch
    from hymd_b200 import _lib
    from hymd_b200.pm import ParticleMesh
    x = torch.zeros((4, 3))
    k0 = ParticleMesh._fingerprint(x)
    _lib.mark_written(x)
    assert ParticleMesh._fingerprint(x) != k0
    a = np.zeros((100, 3))
    k1 = ParticleMesh._fingerprint(a)
    a[50, 1] = 1.0
    assert ParticleMesh._fingerprint(a) != k1


def test_virtual_ranks_host_logic():
    """hymd_b200._world.VirtualRanks: P threads see themselves as P ranks, all-reduce sums over them, an
    exception in one rank surfaces in the caller (no GPU needed: hymd_local_group_id is host code)."""
    import torch
    from hymd_b200 import _world

    def worker(rank):
        w = _world.current()
        assert (w.size, w.rank) == (3, rank)
        return float(w.allreduce(torch.tensor([float(rank + 1)], dtype=torch.float64))[0])

    assert _world.VirtualRanks(3).run(worker) == [6.0, 6.0, 6.0]
    assert _world.current().size == 1

    def failing(rank):
        if rank == 1:
            raise ValueError("rank 1 fails")
        return _world.current().allreduce(torch.zeros(1))

    with pytest.raises(ValueError, match="rank 1 fails"):
        _world.VirtualRanks(3, timeout=5.0).run(failing)

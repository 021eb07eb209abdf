"""Row f4: ``hymd_b200.file_io`` against golden H5MD trees written by the REFERENCE's own ``hymd/file_io.py``.

``h5py`` is not installed in this image.  ``tests/golden/make_file_io_golden.py`` executed the reference's unmodified
``store_static`` / ``store_data`` / ``distribute_input`` over the in-memory stand-in ``tests/fake_h5.py`` and committed
every group, dataset (dtype, shape, values), attribute and log line they produced; here the same inputs go through
``hymd_b200.file_io`` over the same stand-in -- as numpy arrays and as torch tensors (the code path device tensors
take) -- and the trees must agree.  The reference's own assertions (``test/test_file_io.py``,
``test/test_distribute_input.py``) are restated as well.  What stays unpinned is h5py / HDF5 itself."""
import json
import os

import numpy as np
import pytest
import torch

import fake_h5
from hymd_b200 import file_io as fio

HERE = os.path.dirname(os.path.abspath(__file__))
_G = np.load(os.path.join(HERE, "golden", "file_io_golden.npz"))
G = {str(k): _G[f"a{i}"] for i, k in enumerate(_G["__keys__"])}

# values that legitimately differ: who ran it, which program wrote it, how the config object prints itself
NOT_COMPARED = {"/h5md/author@name", "/h5md/creator@version"}
# written from float64 device sums here, from numpy sums in the input dtype there
SUMS = {"/observables/total_momentum/value", "/observables/angular_momentum/value", "/observables/torque/value"}


class Cfg:
    """The attributes file_io reads from the reference's Config."""

    def __init__(self, text, **kw):
        self._text = text
        maps = json.loads(str(G["A|maps"]))
        self.name_to_type_map = maps["name_to_type"]
        self.type_to_name_map = {int(k): v for k, v in maps["type_to_name"].items()}
        self.n_types = maps["n_types"]
        self.n_particles = len(G["in|indices"])
        self.target_temperature = None
        self.initial_energy = None
        self.thermostat_work = 0.0
        self.max_molecule_size = None
        self.__dict__.update(kw)

    def __str__(self):
        return self._text


@pytest.fixture()
def backend():
    fio.set_backend(fake_h5)
    yield fake_h5
    fio.set_backend(None)


def compare(case, h5file):
    mine = fake_h5.tree(h5file)
    want = {k[len(case) + 1:]: v for k, v in G.items() if k.startswith(case + "|/")}
    assert sorted(mine) == sorted(want), set(mine) ^ set(want)
    for path, ref in want.items():
        got = mine[path]
        if path.endswith("/"):
            assert got is None
            continue
        if path in NOT_COMPARED:
            assert np.asarray(got).dtype.kind == ref.dtype.kind
            continue
        got = np.asarray(got)
        assert got.dtype == ref.dtype or (got.dtype.kind in "US" and ref.dtype.kind == got.dtype.kind), (path, got.dtype, ref.dtype)
        assert got.shape == ref.shape, path
        if path in SUMS:
            np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-9, err_msg=path)
        elif got.dtype.kind == "f":
            np.testing.assert_array_equal(got, ref, err_msg=path)
        else:
            assert np.array_equal(got, ref), path


def wrap(x, how):
    if how == "numpy" or not isinstance(x, np.ndarray) or x.dtype.kind not in "fi":
        return x
    return torch.from_numpy(np.ascontiguousarray(x))


@pytest.mark.parametrize("how", ["numpy", "torch"])
def test_case_a_the_references_own_store_data_test(backend, tmp_path, how):
    """``test/test_file_io.py::test_store_static`` / ``test_store_data``: fixture ``molecules_with_solvent``,
    ``charges=True``, ``plumed_out=True``; two frames; the whole tree against the reference's."""
    ind, pos, mol, vel = G["in|indices"], G["in|positions"], G["in|molecules"], G["in|velocities"]
    names, types = G["in|names"], G["in|types"]
    cfg = Cfg(str(G["A|config_str"]), n_steps=100, n_print=1, mass=72.0)
    out = fio.OutDataset(tmp_path, cfg)
    assert out.float_dtype == "float32" and not out.disable_mpio and out.is_open()
    assert out.file.filename.endswith("sim.H5")
    rank_range, flag = fio.distribute_input({"molecules": mol, "indices": ind}, 0, 1, cfg.n_particles, 6)
    assert flag and rank_range == list(range(len(ind)))
    fio.store_static(out, rank_range, names, types, ind, cfg, G["in|bonds_2_atom1"], G["in|bonds_2_atom2"],
                     molecules=mol, charges=True, plumed_out=True)
    # the reference's own assertions
    assert all(k in out.file.keys() for k in ["connectivity", "h5md", "observables", "parameters", "particles"])
    assert all(k in out.file["particles/all"] for k in ["box", "mass", "position", "species"])
    assert all(k in out.file["observables"] for k in [
        "angle_energy", "angular_momentum", "bond_energy", "dihedral_energy", "field_energy", "kinetic_energy",
        "potential_energy", "temperature", "thermostat_work", "torque", "total_energy", "total_momentum",
        "field_q_energy", "plumed_bias"])
    assert "vmd_structure" in out.file["parameters"].keys()
    forces = pos.copy()
    box = np.array([10.0, 10.0, 10.0])
    w = lambda a: wrap(a, how)      # noqa: E731
    fio.store_data(out, 0, 0, w(ind), w(pos), w(vel), w(forces), box, 300., 1., 1., 2., 3., 4., 5., 6., 7., 0.02, cfg,
                   charge_out=True, plumed_out=True)
    assert out.last_log == str(G["A|log"][0])
    for stem in ("positions", "total_energy", "potential_energy", "kinetc_energy", "bond_energy", "angle_energy",
                 "dihedral_energy", "field_energy", "total_momentum", "angular_momentum", "torque", "temperature",
                 "thermostat_work"):
        assert getattr(out, stem + "_step")[0] == 0 and getattr(out, stem + "_time")[0] == 0
    assert out.potential_energy[0] == pytest.approx(20.) and out.kinetc_energy[0] == pytest.approx(1.)
    assert out.temperature[0] == pytest.approx(300.) and out.pressure[0] == pytest.approx(1.)
    assert out.bond_energy[0] == pytest.approx(2.) and out.angle_energy[0] == pytest.approx(3.)
    assert out.dihedral_energy[0] == pytest.approx(4.) and out.field_energy[0] == pytest.approx(5.)
    assert out.field_q_energy[0] == pytest.approx(6.) and out.plumed_bias[0] == pytest.approx(7.)
    # energies as 0-d tensors, the way the device-resident loop hands them over
    e = (lambda x: torch.tensor(x, dtype=torch.float64)) if how == "torch" else (lambda x: x)
    fio.store_data(out, 1, 1, w(ind), w(pos + 0.25), w(2.0 * vel), w(-forces), box, 310., np.arange(18.0), e(1.5),
                   e(2.5), e(0.0), e(4.5), e(5.5), e(0.0), 0.0, 0.02, cfg, charge_out=True, plumed_out=True)
    assert out.last_log == str(G["A|log"][1])
    compare("A", out.file)
    out.flush()
    out.close_file()
    assert not out.is_open()


@pytest.mark.parametrize("how", ["numpy", "torch"])
def test_case_b_everything_on_and_a_shuffled_particle_order(backend, tmp_path, how):
    """Velocities and forces out, float64 output, per-particle charges / dielectrics, no molecules, ``n_print`` 10,
    per-particle log columns of an NVT run; ``store_data`` gets the rows in the shuffled order a domain decomposition
    leaves behind and must put them back by global index (h5py takes point selections in increasing order only:
    the stand-in enforces it)."""
    ind, pos, vel = G["in|indices"], G["in|positions"], G["in|velocities"]
    perm = G["in|perm"]
    cfg = Cfg(str(G["B|config_str"]), n_steps=50, n_print=10, mass=72.0, target_temperature=323.0,
              initial_energy=123.5, thermostat_work=7.25)
    out = fio.OutDataset(tmp_path, cfg, double_out=True)
    assert out.float_dtype == "float64"
    fio.store_static(out, list(range(len(ind))), G["in|names"], G["in|types"], ind, cfg, np.zeros(0, dtype=int),
                     np.zeros(0, dtype=int), molecules=None, velocity_out=True, force_out=True,
                     charges=G["in|charges"], dielectrics=G["in|dielectrics"])
    assert fio.n_output_frames(cfg) == 6 and out.positions.shape == (6, 45, 3)
    forces = pos.copy()
    w = lambda a: wrap(a, how)      # noqa: E731
    ip = w(np.ascontiguousarray(ind[perm]))
    fio.store_data(out, 20, 2, ip, w(pos[perm]), w(vel[perm]), w(forces[perm] * 3.0), np.array([9.5, 10.0, 10.5]),
                   323.0, np.linspace(-1, 1, 18), 11.0, 12.0, 13.0, 14.0, 15.0, 16.0, 0.0, 0.03, cfg,
                   velocity_out=True, force_out=True, charge_out=True, dump_per_particle=True)
    assert out.last_log == str(G["B|log"][0])
    compare("B", out.file)
    # the cached permutation follows the indices object: new content, new order
    perm2 = perm[::-1].copy()
    fio.store_data(out, 30, 3, w(np.ascontiguousarray(ind[perm2])), w(pos[perm2]), w(vel[perm2]), w(forces[perm2]),
                   np.array([9.5, 10.0, 10.5]), 323.0, 0.0, 11.0, 12.0, 13.0, 14.0, 15.0, 16.0, 0.0, 0.03, cfg,
                   velocity_out=True, force_out=True, charge_out=True)
    assert np.array_equal(out.positions[3], pos) and np.array_equal(out.velocities[3], vel)
    if how == "torch":          # in-place change of the same tensor (version counter)
        ip.copy_(torch.from_numpy(np.ascontiguousarray(ind[perm2])))
        fio.store_data(out, 40, 4, ip, w(pos[perm2]), w(vel[perm2]), w(forces[perm2]), np.array([9.5, 10.0, 10.5]),
                       323.0, 0.0, 11.0, 12.0, 13.0, 14.0, 15.0, 16.0, 0.0, 0.03, cfg, velocity_out=True,
                       force_out=True, charge_out=True)
        assert np.array_equal(out.positions[4], pos)


def test_distribute_input_matches_the_reference_for_every_rank_count():
    ind = np.arange(0, 10000)
    mol = np.zeros_like(ind)
    mol[400:450] = 1
    mol[450:] = np.arange(2, 9552)
    mol2 = np.concatenate([np.repeat(np.arange(300), 12), 300 + np.arange(2400)])
    ind2 = np.arange(len(mol2))
    seen = 0
    for key, want in G.items():
        if not key.startswith("D|"):
            continue
        _, kind, size, rank = key.split("|")
        size, rank = int(size), int(rank)
        if kind == "mol":
            rr, flag = fio.distribute_input({"indices": ind, "molecules": mol}, rank, size, len(ind), max_molecule_size=1000)
        elif kind == "nomol":
            rr, flag = fio.distribute_input({"indices": ind}, rank, size, None)
        else:
            rr, flag = fio.distribute_input({"indices": ind2, "molecules": mol2}, rank, size, len(ind2))
        assert [rr[0], rr[-1] + 1, int(flag)] == list(want), key
        assert rr == list(range(rr[0], rr[-1] + 1))
        seen += 1
    assert seen == 2 * (1 + 2 + 3 + 5 + 8) + (2 + 4 + 7)
    # the properties test/test_distribute_input.py asserts: contiguous cover, breaks only between molecules
    for size in (5, 9, 11, 14, 19, 25):
        ranges = [fio.distribute_input({"indices": ind2, "molecules": mol2}, r, size, len(ind2))[0] for r in range(size)]
        assert np.array_equal(np.concatenate(ranges), ind2)
        for r in ranges[1:]:
            assert mol2[r[0] - 1] != mol2[r[0]]


def test_time_dependent_element_and_output_modes(backend, tmp_path):
    """``test_OutDataset`` / ``test_setup_time_dependent_element`` of the reference."""
    cfg = Cfg("x", n_steps=10, n_print=5, mass=72.0)
    out = fio.OutDataset(tmp_path, cfg)
    g = out.file.create_group("/test")
    group, step, time, value = fio.setup_time_dependent_element("position", g, 1, (cfg.n_particles, 3), "float32",
                                                                units="nm")
    assert group.name == "/test/position" and value.shape == (1, cfg.n_particles, 3)
    assert step.dtype == np.int32 and time.dtype == np.float32 and value.dtype == np.float32
    assert time.attrs["unit"] == "ps" and value.attrs["unit"] == "nm"
    out.close_file()
    assert not out.file
    out = fio.OutDataset(tmp_path, cfg, disable_mpio=True)
    assert out.disable_mpio and out.file.filename.endswith("sim.hdf5-     0-of-     1")
    out.close_file()


def test_several_ranks_write_their_rows_and_their_bonds_at_global_offsets(backend, tmp_path):
    """Two ranks (threads of this process, ``VirtualRanks``): one file per rank named like the reference's
    ``disable_mpio`` mode, bonds at the offset of the lower ranks' bond counts, momenta summed over the ranks."""
    from hymd_b200._world import VirtualRanks
    ind, pos, vel, mol = G["in|indices"], G["in|positions"], G["in|velocities"], G["in|molecules"]
    names, types = G["in|names"], G["in|types"]
    b1, b2 = G["in|bonds_2_atom1"], G["in|bonds_2_atom2"]
    cfg = Cfg(str(G["A|config_str"]), n_steps=100, n_print=1, mass=72.0)
    files = [None, None]

    def worker(rank):
        rr, _ = fio.distribute_input({"molecules": mol, "indices": ind}, rank, 2, cfg.n_particles, 6)
        lo, hi = rr[0], rr[-1] + 1
        keep = (b1 >= lo) & (b1 < hi)          # bonds of the molecules homed here, in local numbering
        out = fio.OutDataset(tmp_path, cfg)
        fio.store_static(out, rr, names[lo:hi], types[lo:hi], ind[lo:hi], cfg, b1[keep] - lo, b2[keep] - lo,
                         molecules=mol[lo:hi], charges=True, plumed_out=True)
        fio.store_data(out, 0, 0, ind[lo:hi], pos[lo:hi], vel[lo:hi], pos[lo:hi].copy(), np.array([10.0, 10.0, 10.0]),
                       300., 1., 1., 2., 3., 4., 5., 6., 7., 0.02, cfg, charge_out=True, plumed_out=True)
        files[rank] = out
        return lo, hi

    spans = VirtualRanks(2).run(worker)
    assert spans[0][1] == spans[1][0] and mol[spans[0][1] - 1] != mol[spans[0][1]]
    assert files[0].file.filename.endswith("sim.hdf5-     0-of-     2")
    assert files[1].file.filename.endswith("sim.hdf5-     1-of-     2")
    ref_pos = G["A|/particles/all/position/value"][0]
    ref_from, ref_to = G["A|/parameters/vmd_structure/bond_from"], G["A|/parameters/vmd_structure/bond_to"]
    got_pos = np.zeros_like(ref_pos)
    got_from, got_to = np.zeros_like(ref_from), np.zeros_like(ref_to)
    for r, (lo, hi) in enumerate(spans):
        f = files[r].file
        got_pos[lo:hi] = f["particles/all/position/value"][0, lo:hi]
        assert not f["particles/all/position/value"][0, :lo].any() and not f["particles/all/position/value"][0, hi:].any()
        got_from += f["parameters/vmd_structure/bond_from"][:]
        got_to += f["parameters/vmd_structure/bond_to"][:]
        # sums over ALL ranks in every file
        np.testing.assert_allclose(f["observables/total_momentum/value"][0], G["A|/observables/total_momentum/value"][0],
                                   rtol=2e-6, atol=1e-9)
        # (forces = positions in this fixture: r x f vanishes up to the order of summation)
        np.testing.assert_allclose(f["observables/torque/value"][0], G["A|/observables/torque/value"][0], rtol=2e-6,
                                   atol=1e-8)
        np.testing.assert_allclose(f["observables/angular_momentum/value"][0],
                                   G["A|/observables/angular_momentum/value"][0], rtol=2e-6, atol=1e-8)
    assert np.array_equal(got_pos, ref_pos)
    assert np.array_equal(got_from, ref_from) and np.array_equal(got_to, ref_to)


def test_read_input_follows_main_py(backend):
    """``main.py:72-125`` over an in-memory input file: last frame of coordinates / velocities, this rank's rows,
    box attribute, optional datasets."""
    ind, pos, vel, mol = G["in|indices"], G["in|positions"], G["in|velocities"], G["in|molecules"]
    f = fake_h5.File("in.h5", "w")
    f.create_dataset("indices", data=ind)
    f.create_dataset("coordinates", data=np.stack([pos + 1.0, pos]))
    f.create_dataset("velocities", data=np.stack([vel * 0.0, vel]))
    f.create_dataset("names", data=G["in|names"])
    f.create_dataset("types", data=G["in|types"])
    f.create_dataset("molecules", data=mol)
    f.create_dataset("bonds", data=G["in|bonds"])
    f.create_dataset("charge", data=G["in|charges"])
    f.attrs["box"] = np.array([10.0, 11.0, 12.0])
    cfg = Cfg("x", box_size=None)
    got = [fio.read_input(f, cfg, dtype=np.float32, rank=r, size=3) for r in range(3)]
    assert np.array_equal(np.concatenate([g["indices"] for g in got]), ind)
    assert np.array_equal(np.concatenate([g["positions"] for g in got]), pos.astype(np.float32))
    assert np.array_equal(np.concatenate([g["velocities"] for g in got]), vel.astype(np.float32))
    assert np.array_equal(np.concatenate([g["charges"] for g in got]), G["in|charges"])
    assert np.array_equal(np.concatenate([g["bonds"] for g in got]), G["in|bonds"])
    assert got[0]["positions"].dtype == np.float32 and all(g["molecules_flag"] for g in got)
    assert np.array_equal(cfg.box_size, [10.0, 11.0, 12.0])
    for a, b in zip(got[:-1], got[1:]):
        assert mol[a["rank_range"][-1]] != mol[b["rank_range"][0]]
    # without molecules, velocities and box attribute
    g2 = fake_h5.File("in2.h5", "w")
    g2.create_dataset("indices", data=ind)
    g2.create_dataset("coordinates", data=pos[None])
    g2.create_dataset("names", data=G["in|names"])
    cfg2 = Cfg("x", box_size=np.array([5.0, 5.0, 5.0]))
    r = fio.read_input(g2, cfg2, dtype=np.float64, rank=0, size=1, topol={})
    assert not r["molecules_flag"] and not r["velocities"].any() and r["types"] is None and r["charges"] is None
    assert r["bonds"] is None and r["molecules"] == [] and r["positions"].dtype == np.float64
    with pytest.raises(ValueError):
        fio.read_input(g2, Cfg("x", box_size=None), rank=0, size=1)


def test_without_h5py_the_layer_says_so(tmp_path):
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py present")
    except ImportError:
        pass
    from hymd_b200._lib import HymdError
    fio.set_backend(None)
    with pytest.raises(HymdError, match="h5py"):
        fio.OutDataset(tmp_path, Cfg("x"))


@pytest.mark.parametrize("nproc", [2, 3])
def test_file_io_over_process_ranks(nproc):
    """torch.distributed (gloo) ranks instead of threads: tests/gloo_file_io_worker.py."""
    from conftest import ROOT
    from test_gloo_slabs import _torchrun
    r = _torchrun(nproc, os.path.join(ROOT, "tests", "gloo_file_io_worker.py"), [], 29641 + nproc)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "OK" in r.stdout

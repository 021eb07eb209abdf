"""Golden H5MD trees produced by EXECUTING the reference's own ``hymd/file_io.py`` in the build container.

    python tests/golden/make_file_io_golden.py        (needs /root/reference)

``h5py`` is not installed here, so the reference's unmodified ``store_static`` / ``store_data`` / ``distribute_input``
run over ``tests/fake_h5.py`` (registered as ``sys.modules["h5py"]``) and a one-rank MPI stub (``ref_loader``).  What
they leave in the in-memory file -- every group, dataset (dtype, shape, values) and attribute -- is flattened into
``tests/golden/file_io_golden.npz`` together with the inputs, the log line ``store_data`` emits and the rank ranges
``distribute_input`` returns; ``tests/test_file_io.py`` replays the same inputs through ``hymd_b200.file_io``.

Cases:
* ``A``  the reference's own ``test_store_data`` (``test/test_file_io.py:187-330``): ``molecules_with_solvent``
  fixture, ``charges=True``, ``plumed_out=True``, two frames;
* ``B``  velocities + forces out, double precision, per-particle charge and dielectric arrays, no molecules,
  ``n_print = 10``, ``dump_per_particle``, an NVT run with thermostat work (the ``H~`` column);
* ``D*`` ``distribute_input`` on the reference's ``test_distribute_input`` system for 1, 2, 3, 5 and 8 ranks, and on
  a system without molecules.
"""
import collections
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import fake_h5  # noqa: E402
import ref_loader as rl  # noqa: E402


class Capture:
    def __init__(self):
        self.lines = []

    def log(self, level, msg):
        self.lines.append(msg)


def flatten(prefix, h5file, out, keys):
    for path, val in fake_h5.tree(h5file).items():
        keys.append(prefix + "|" + path)
        if val is None:
            out.append(np.zeros(0))
        elif isinstance(val, str):
            out.append(np.array(val))
        else:
            out.append(np.asarray(val))


def main():
    rl.install_stubs()
    sys.modules["h5py"] = fake_h5
    fio = rl.ref("file_io")
    ip = rl.ref("input_parser")
    force = rl.ref("force")
    cap = Capture()
    fio.Logger.rank0 = cap
    comm = sys.modules["mpi4py"].MPI.COMM_WORLD

    keys, vals = [], []

    def put(k, v):
        keys.append(k)
        vals.append(np.asarray(v))

    indices, positions, molecules, velocities, bonds, names, types_ = rl.conftest_fixture("molecules_with_solvent")
    Bond = collections.namedtuple("Bond", ["atom_1", "atom_2", "equilibrium", "strength"])
    cbonds = tuple(Bond(a, b, 0.27, 10000) for a, b in (("A", "A"), ("A", "B"), ("A", "C"), ("B", "B"), ("B", "C")))
    for k, v in (("indices", indices), ("positions", positions), ("molecules", molecules), ("velocities", velocities),
                 ("bonds", bonds), ("names", names), ("types", types_)):
        put("in|" + k, v)

    def make_config(**kw):
        config = ip.Config(time_step=0.03, box_size=np.array([10, 10, 10], dtype=np.float64), mesh_size=[5, 5, 5],
                           sigma=0.5, kappa=0.05, n_particles=len(indices),
                           target_pressure=types.SimpleNamespace(P_L=None, P_N=None), **kw)
        config.bonds = cbonds
        return ip._setup_type_to_name_map(config, names, types_)

    # ---- case A: the reference's test_store_data, two frames -------------------------------------------------------
    config = make_config(n_steps=100, n_print=1, mass=72.0)
    in_file = {"molecules": molecules, "indices": indices}
    rank_range, _ = fio.distribute_input(in_file, 0, 1, config.n_particles, 6, comm=comm)
    prep = force.prepare_bonds(molecules[rank_range], names[rank_range], bonds[rank_range], indices[rank_range], config)
    b2a, b2b = np.asarray(prep[0]), np.asarray(prep[1])
    put("in|bonds_2_atom1", b2a)
    put("in|bonds_2_atom2", b2b)
    out = fio.OutDataset("/tmp", config, comm=comm)
    fio.store_static(out, rank_range, names, types_, indices, config, b2a, b2b, molecules=molecules, charges=True,
                     plumed_out=True)
    forces = np.copy(positions)
    fio.store_data(out, 0, 0, indices, positions, velocities, forces, config.box_size, 300., 1., 1., 2., 3., 4., 5., 6.,
                   7., 0.02, config, charge_out=True, plumed_out=True)
    fio.store_data(out, 1, 1, indices, positions + 0.25, 2.0 * velocities, -forces, config.box_size, 310.,
                   np.arange(18.0), 1.5, 2.5, 0.0, 4.5, 5.5, 0.0, 0.0, 0.02, config, charge_out=True, plumed_out=True)
    flatten("A", out.file, vals, keys)
    put("A|log", np.array(cap.lines))
    put("A|config_str", np.array(str(config)))
    put("A|maps", np.array(json.dumps({"name_to_type": {k: int(v) for k, v in config.name_to_type_map.items()},
                                       "type_to_name": {str(int(k)): v for k, v in config.type_to_name_map.items()},
                                       "n_types": int(config.n_types)})))
    cap.lines = []

    # ---- case B: everything switched on, shuffled local order ------------------------------------------------------
    config = make_config(n_steps=50, n_print=10, mass=72.0, target_temperature=323.0)
    config.initial_energy = 123.5
    config.thermostat_work = 7.25
    rng = np.random.default_rng(77)
    perm = rng.permutation(len(indices))            # the order domain_decomposition leaves behind
    charges = rng.normal(size=len(indices)).astype(np.float32)
    dielectrics = (1.0 + 79.0 * rng.random(len(indices))).astype(np.float32)
    put("in|perm", perm)
    put("in|charges", charges)
    put("in|dielectrics", dielectrics)
    out = fio.OutDataset("/tmp", config, double_out=True, comm=comm)
    # (store_static runs at start-up, on the order of the input file: its ``charge[indices] = charges`` is a point
    # selection, which h5py only takes in increasing order; store_data sorts for itself)
    fio.store_static(out, rank_range, names, types_, indices, config, np.zeros(0, dtype=int),
                     np.zeros(0, dtype=int), molecules=None, velocity_out=True, force_out=True, charges=charges,
                     dielectrics=dielectrics)
    fio.store_data(out, 20, 2, indices[perm], positions[perm], velocities[perm], forces[perm] * 3.0,
                   np.array([9.5, 10.0, 10.5]), 323.0, np.linspace(-1, 1, 18), 11.0, 12.0, 13.0, 14.0, 15.0, 16.0, 0.0,
                   0.03, config, velocity_out=True, force_out=True, charge_out=True, dump_per_particle=True)
    flatten("B", out.file, vals, keys)
    put("B|log", np.array(cap.lines))
    put("B|config_str", np.array(str(config)))
    cap.lines = []

    # ---- distribute_input -----------------------------------------------------------------------------------------
    ind = np.arange(0, 10000)
    mol = np.zeros_like(ind)
    mol[400:450] = 1
    mol[450:] = np.arange(2, 9552)
    for size in (1, 2, 3, 5, 8):
        for rank in range(size):
            rr, flag = fio.distribute_input({"indices": ind, "molecules": mol}, rank, size, len(ind),
                                            max_molecule_size=1000, comm=comm)
            put(f"D|mol|{size}|{rank}", [rr[0], rr[-1] + 1, int(flag)])
            rr, flag = fio.distribute_input({"indices": ind}, rank, size, len(ind), comm=comm)
            put(f"D|nomol|{size}|{rank}", [rr[0], rr[-1] + 1, int(flag)])
    # chains of 12 beads + solvent, default max_molecule_size
    mol2 = np.concatenate([np.repeat(np.arange(300), 12), 300 + np.arange(2400)])
    ind2 = np.arange(len(mol2))
    for size in (2, 4, 7):
        for rank in range(size):
            rr, flag = fio.distribute_input({"indices": ind2, "molecules": mol2}, rank, size, len(ind2), comm=comm)
            put(f"D|chains|{size}|{rank}", [rr[0], rr[-1] + 1, int(flag)])

    path = os.path.join(HERE, "file_io_golden.npz")
    np.savez_compressed(path, __keys__=np.array(keys), **{f"a{i}": v for i, v in enumerate(vals)})
    print(path, len(keys), "entries")


if __name__ == "__main__":
    main()

"""HDF5 input and H5MD output for a run whose particle data lives on the GPU.

Mirrors ``hymd/file_io.py`` of the reference (HyMD v2.2.0) -- ``OutDataset`` (``file_io.py:12-86``),
``setup_time_dependent_element`` (``:89-119``), ``store_static`` (``:122-517``), ``store_data`` (``:528-777``),
``distribute_input`` (``:780-874``) -- with the same names, positional arguments, file layout, dataset dtypes / shapes,
attributes and log line, plus ``read_input`` for the start-up reads of ``main.py:72-125``.  The files are the ones VMD's
h5md plugin and HyMD's own tools expect; a trajectory written here continues to work as an input file there.

What is different, and why:

* **the per-particle arrays are device tensors.**  ``store_data`` takes ``positions`` / ``velocities`` / ``forces`` /
  ``indices`` as torch tensors on the GPU (numpy arrays work too): the sums behind ``total_momentum``,
  ``angular_momentum`` and ``torque`` (``file_io.py:682-689``) are evaluated on the device in float64 and cross PCIe
  as 9 numbers; each per-particle array is permuted into global index order **on the device** (the reference's
  ``argsort(indices)``, ``file_io.py:667-673``; the permutation is cached until ``indices`` changes) and copied once
  through a persistent pinned staging buffer, so that the host hands h5py one contiguous block per dataset;
* **no MPI-IO.**  One process per GPU and a serial HDF5 library: with one rank the file is ``sim.H5`` as in the
  reference; with several ranks every rank writes its own rows into its own full-size file, named as the reference's
  ``disable_mpio`` mode names them (``file_io.py:50-56``).  ``comm`` arguments are accepted and ignored: the rank
  layout is ``hymd_b200._world.current()``;
* **h5py is imported on first use.**  It is not a dependency of the field-force path and is absent from the image this
  was built in: without it ``OutDataset`` / ``read_input`` raise ``HymdError`` (no silent alternative format).  The
  layer is exercised over an in-memory stand-in (``set_backend``; ``tests/fake_h5.py``) against golden trees written by
  the reference's own ``file_io.py`` running over the same stand-in (``tests/golden/make_file_io_golden.py``).
"""
from __future__ import annotations

import getpass
import logging
import os

import numpy as np
import torch

from . import __version__, _world
from ._lib import HymdError

_backend = None
log_rank0 = logging.getLogger("hymd_b200.rank_0")


def set_backend(module):
    """Use ``module`` (anything with h5py's ``File``) instead of importing h5py; ``None`` restores the default."""
    global _backend
    _backend = module


def _h5():
    if _backend is not None:
        return _backend
    try:
        import h5py
    except ImportError as e:
        raise HymdError("hymd_b200.file_io needs h5py for HDF5 input / H5MD output and it is not installed "
                        "(there is no alternative output format)") from e
    return h5py


def _host(x):
    """numpy view / copy of a small array-like (tensor, DeviceScalar, list, scalar)."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    if hasattr(x, "item") and not isinstance(x, np.ndarray) and not isinstance(x, np.generic):
        return np.asarray(x.item())
    return np.asarray(x)


def _scalar(x):
    return float(_host(x).reshape(-1)[0]) if np.size(_host(x)) == 1 else _host(x).astype(np.float64)


class OutDataset:
    """HDF5 output handle (``file_io.py:12-86``): ``.file``, ``.float_dtype``, ``.config``, ``.disable_mpio``."""

    def __init__(self, dest_directory, config, double_out=False, disable_mpio=False, comm=None):
        self.config = config
        self.float_dtype = "float64" if double_out else "float32"
        world = _world.current()
        # several ranks without MPI-IO = the reference's one-file-per-rank mode
        self.disable_mpio = bool(disable_mpio)
        self.per_rank = self.disable_mpio or world.size > 1
        if self.per_rank:
            name = f"sim.hdf5-{world.rank:6d}-of-{world.size:6d}"
        else:
            name = "sim.H5"
        self.file = _h5().File(os.path.join(str(dest_directory), name), "w")
        self._stage = {}          # role -> pinned staging tensor
        self._order = None        # (key, sorted global indices (numpy), permutation (tensor), contiguous?)
        self.last_log = None

    def is_open(self, comm=None):
        return bool(self.file)

    def close_file(self, comm=None):
        self.file.close()

    def flush(self):
        self.file.flush()


def setup_time_dependent_element(name, parent_group, n_frames, shape, dtype, units=None):
    """H5MD time-dependent element (``file_io.py:89-119``): group ``name`` with ``step (n_frames,) int32``,
    ``time (n_frames,) float32`` and ``value (n_frames, *shape) dtype``; ``unit`` attributes when given."""
    group = parent_group.create_group(name)
    step = group.create_dataset("step", (n_frames,), "int32")
    time = group.create_dataset("time", (n_frames,), "float32")
    value = group.create_dataset("value", (n_frames, *shape), dtype)
    if units is not None:
        value.attrs["unit"] = units
        time.attrs["unit"] = "ps"
    return group, step, time, value


def n_output_frames(config):
    """Frames the output file holds (``file_io.py:225-231``)."""
    n_frames = config.n_steps // config.n_print
    if (config.n_steps - 1) % config.n_print != 0:
        n_frames += 1
    if config.n_steps % config.n_print == 1:
        n_frames += 1
    if n_frames == config.n_steps:
        n_frames += 1
    return n_frames


# (attribute stem on the OutDataset, group name, trailing shape, unit, forced dtype or None) in the reference's order
_OBSERVABLES = [
    ("total_energy", "total_energy", (1,), "kJ mol-1", None),
    ("kinetc_energy", "kinetic_energy", (1,), "kJ mol-1", None),       # (sic: the reference's attribute name)
    ("potential_energy", "potential_energy", (1,), "kJ mol-1", None),
    ("bond_energy", "bond_energy", (1,), "kJ mol-1", None),
    ("angle_energy", "angle_energy", (1,), "kJ mol-1", None),
    ("dihedral_energy", "dihedral_energy", (1,), "kJ mol-1", None),
    ("field_energy", "field_energy", (1,), "kJ mol-1", None),
    ("field_q_energy", "field_q_energy", (1,), "kJ mol-1", None),      # only with charges
    ("plumed_bias", "plumed_bias", (1,), "kJ mol-1", None),            # only with plumed_out
    ("total_momentum", "total_momentum", (3,), "nm g ps-1 mol-1", None),
    ("angular_momentum", "angular_momentum", (3,), "nm+2 g ps-1 mol-1", None),
    ("torque", "torque", (3,), "kJ nm+2 mol-1", None),
    ("temperature", "temperature", (3,), "K", None),
    ("thermostat_work", "thermostat_work", (1,), "kJ mol-1", "float32"),
    ("pressure", "pressure", (18,), "Bar", "float32"),
]


def _bind(h5md, stem, parent, name, n_frames, shape, dtype, units):
    _, step, time, value = setup_time_dependent_element(name, parent, n_frames, shape, dtype, units=units)
    setattr(h5md, stem + "_step", step)
    setattr(h5md, stem + "_time", time)
    setattr(h5md, stem, value)


def _rank_offset(n_local):
    """(sum over lower ranks, total) of a per-rank count."""
    world = _world.current()
    if world.size == 1:
        return 0, int(n_local)
    v = torch.zeros(world.size, dtype=torch.float64)
    v[world.rank] = float(n_local)
    v = world.allreduce(v).cpu()
    return int(v[:world.rank].sum().item()), int(v.sum().item())


def store_static(h5md, rank_range, names, types, indices, config, bonds_2_atom1, bonds_2_atom2, molecules=None,
                 velocity_out=False, force_out=False, charges=False, dielectrics=False, plumed_out=False, comm=None):
    """Everything that does not change during the run (``file_io.py:122-517``): the H5MD skeleton (``/h5md``,
    ``/observables``, ``/connectivity``, ``/parameters``, ``/particles/all``), masses, charges, species, the
    time-dependent elements ``store_data`` fills, and the ``vmd_structure`` group (names, types, residue ids, bonds).
    Called once at start-up, in the particle order of the input file (``indices`` increasing on every rank)."""
    dtype = h5md.float_dtype
    f = h5md.file
    h5md.h5md_group = f.create_group("/h5md")
    h5md.observables = f.create_group("/observables")
    h5md.connectivity = f.create_group("/connectivity")
    h5md.parameters = f.create_group("/parameters")
    h5md.h5md_group.attrs["version"] = np.array([1, 1], dtype=int)
    h5md.h5md_group.create_group("author").attrs["name"] = np.bytes_(getpass.getuser())
    creator = h5md.h5md_group.create_group("creator")
    creator.attrs["name"] = np.bytes_("Hylleraas MD")
    creator.attrs["version"] = np.bytes_(f"hymd_b200 {__version__}")

    N = int(config.n_particles)
    idx = _host(indices).astype(np.int64).reshape(-1)
    order = np.argsort(idx, kind="stable")
    idx_sorted = idx[order]
    h5md.particles_group = f.create_group("/particles")
    h5md.all_particles = h5md.particles_group.create_group("all")
    mass = h5md.all_particles.create_dataset("mass", (N,), dtype)
    mass[...] = config.mass
    for label, values in (("charge", charges), ("dielectric", dielectrics)):
        if values is not False:
            dset = h5md.all_particles.create_dataset(label, (N,), dtype="float32")
            # (``charges=True`` -- what the reference's own test passes -- broadcasts 1.0 like numpy does there)
            v = _host(values)
            dset[idx_sorted] = v[order] if v.ndim else v
    box = h5md.all_particles.create_group("box")
    box.attrs["dimension"] = 3
    box.attrs["boundary"] = np.array([b"periodic"] * 3, dtype="S8")

    n_frames = n_output_frames(config)
    species = h5md.all_particles.create_dataset("species", (N,), dtype="i")
    _bind(h5md, "positions", h5md.all_particles, "position", n_frames, (N, 3), dtype, "nm")
    if velocity_out:
        _bind(h5md, "velocities", h5md.all_particles, "velocity", n_frames, (N, 3), dtype, "nm ps-1")
    if force_out:
        _bind(h5md, "forces", h5md.all_particles, "force", n_frames, (N, 3), dtype, "kJ mol-1 nm-1")
    for stem, name, shape, unit, forced in _OBSERVABLES:
        if (stem == "field_q_energy" and charges is False) or (stem == "plumed_bias" and plumed_out is False):
            continue
        _bind(h5md, stem, h5md.observables, name, n_frames, shape, forced or dtype, unit)
    _, h5md.box_step, h5md.box_time, h5md.box_value = setup_time_dependent_element(
        "edges", box, n_frames, (3, 3), "float32", units="nm")

    # species of every local particle from its name (file_io.py:466-468), one block write
    nm = _host(names).reshape(-1)
    uniq, inv = np.unique(nm, return_inverse=True)
    to_type = np.array([config.name_to_type_map[u.decode("utf-8") if isinstance(u, bytes) else str(u)] for u in uniq],
                       dtype=np.int32)
    if len(idx):
        species[idx_sorted] = to_type[inv][order]

    h5md.parameters.attrs["config.toml"] = np.bytes_(str(config))
    vmd = h5md.parameters.create_group("vmd_structure")
    vmd.create_dataset("indexOfSpecies", (config.n_types,), "i")[:] = np.arange(config.n_types)
    # the VMD h5md plugin reads at most 16 characters of a name / type
    name_dataset = vmd.create_dataset("name", (config.n_types,), "S16")
    type_dataset = vmd.create_dataset("type", (config.n_types,), "S16")
    resid = vmd.create_dataset("resid", (N,), "i") if molecules is not None else None
    for t, n in config.type_to_name_map.items():
        name_dataset[t] = np.bytes_(n[:16])
        type_dataset[t] = np.bytes_("solvent" if n == "W" else "membrane")

    a1 = _host(bonds_2_atom1).astype(np.int64).reshape(-1)
    a2 = _host(bonds_2_atom2).astype(np.int64).reshape(-1)
    start, total = _rank_offset(len(a1))
    bonds_from = vmd.create_dataset("bond_from", (total,), "i")
    bonds_to = vmd.create_dataset("bond_to", (total,), "i")
    if len(a1):
        bonds_from[start:start + len(a1)] = idx[a1] + 1        # VMD counts from one
        bonds_to[start:start + len(a1)] = idx[a2] + 1
    if resid is not None and len(idx):
        resid[idx_sorted] = _host(molecules).reshape(-1)[order]


def _global_order(h5md, indices):
    """Permutation of the local rows into increasing global index, cached while ``indices`` is the same object with
    the same content version (domain decomposition replaces it)."""
    if isinstance(indices, torch.Tensor):
        key = ("t", indices.data_ptr(), indices._version, tuple(indices.shape))
    else:
        a = np.asarray(indices)
        key = ("n", id(indices), a.ctypes.data, a.shape, hash(a.tobytes()))
    if h5md._order is not None and h5md._order[0] == key:
        return h5md._order[1:]
    t = indices if isinstance(indices, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(indices))
    t = t.reshape(-1).long()
    sorted_idx, perm = torch.sort(t, stable=True)
    host_idx = sorted_idx.cpu().numpy()
    contiguous = len(host_idx) > 0 and int(host_idx[-1] - host_idx[0]) + 1 == len(host_idx)
    h5md._order = (key, host_idx, perm, contiguous, indices)      # (keeps ``indices`` alive: its id is in the key)
    return h5md._order[1:]


def _rows_to_host(h5md, role, x, perm, np_dtype):
    """``x[perm]`` as a host array of the dataset dtype: gather on the device, one copy through pinned memory."""
    if not isinstance(x, torch.Tensor):
        return np.asarray(x)[perm.cpu().numpy()].astype(np_dtype, copy=False)
    g = x.detach().index_select(0, perm.to(x.device))
    tdt = torch.float64 if np.dtype(np_dtype) == np.float64 else torch.float32
    g = g.to(tdt)
    if not g.is_cuda:
        return g.numpy()
    buf = h5md._stage.get(role)
    if buf is None or buf.shape != g.shape or buf.dtype != g.dtype:
        buf = torch.empty(g.shape, dtype=g.dtype, pin_memory=True)
        h5md._stage[role] = buf
    buf.copy_(g, non_blocking=True)
    torch.cuda.current_stream(g.device).synchronize()
    return buf.numpy()


def _write_rows(dset, frame, host_idx, contiguous, rows):
    if len(host_idx) == 0:
        return
    if contiguous:      # a hyperslab, not a point selection
        dset[frame, int(host_idx[0]):int(host_idx[0]) + len(host_idx)] = rows
    else:
        dset[frame, host_idx] = rows


def _as_tensor(x):
    return x.detach() if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


def store_data(h5md, step, frame, indices, positions, velocities, forces, box_size, temperature, pressure,
               kinetic_energy, bond2_energy, bond3_energy, bond4_energy, field_energy, field_q_energy, plumed_bias,
               time_step, config, velocity_out=False, force_out=False, charge_out=False, plumed_out=False,
               dump_per_particle=False, comm=None):
    """One output frame (``file_io.py:528-777``): step / time of every element, positions (velocities, forces) in global
    index order, energies, momenta, torque, temperature, pressure, box, thermostat work; and the log line."""
    stems = ["positions", "total_energy", "potential_energy", "kinetc_energy", "bond_energy", "angle_energy",
             "dihedral_energy", "field_energy", "total_momentum", "angular_momentum", "torque", "temperature",
             "pressure", "box", "thermostat_work"]
    stems += ["velocities"] * bool(velocity_out) + ["forces"] * bool(force_out) + \
        ["field_q_energy"] * bool(charge_out) + ["plumed_bias"] * bool(plumed_out)
    for stem in stems:
        getattr(h5md, stem + "_step")[frame] = step
        getattr(h5md, stem + "_time")[frame] = step * time_step

    host_idx, perm, contiguous, _ = _global_order(h5md, indices)
    out_dtype = np.dtype(h5md.float_dtype)
    _write_rows(h5md.positions, frame, host_idx, contiguous, _rows_to_host(h5md, "x", positions, perm, out_dtype))
    if velocity_out:
        _write_rows(h5md.velocities, frame, host_idx, contiguous, _rows_to_host(h5md, "v", velocities, perm, out_dtype))
    if force_out:
        _write_rows(h5md.forces, frame, host_idx, contiguous, _rows_to_host(h5md, "f", forces, perm, out_dtype))

    # sum v, sum r x v, sum r x f on the device in float64; one all-reduce, one read-back of 9 numbers
    x, v, fr = (_as_tensor(a).double() for a in (positions, velocities, forces))
    if fr.device != x.device:
        fr = fr.to(x.device)
    sums = torch.cat([v.sum(0), torch.linalg.cross(x, v).sum(0), torch.linalg.cross(x, fr).sum(0)]) \
        if x.shape[0] else torch.zeros(9, dtype=torch.float64, device=x.device)
    sums = float(config.mass) * _world.current().allreduce(sums).cpu().numpy()
    total_momentum, angular_momentum, torque = sums[0:3], sums[3:6], sums[6:9]

    kinetic_energy, bond2_energy, bond3_energy, bond4_energy, field_energy, field_q_energy, plumed_bias = (
        _scalar(e) for e in (kinetic_energy, bond2_energy, bond3_energy, bond4_energy, field_energy, field_q_energy,
                             plumed_bias))
    temperature = _scalar(temperature)
    thermostat_work = _scalar(getattr(config, "thermostat_work", 0.0))
    potential_energy = bond2_energy + bond3_energy + bond4_energy + field_energy + field_q_energy
    total_energy = kinetic_energy + potential_energy
    if charge_out:
        h5md.field_q_energy[frame] = field_q_energy
    if plumed_out:
        h5md.plumed_bias[frame] = plumed_bias
    h5md.total_energy[frame] = total_energy
    h5md.potential_energy[frame] = potential_energy
    h5md.kinetc_energy[frame] = kinetic_energy
    h5md.bond_energy[frame] = bond2_energy
    h5md.angle_energy[frame] = bond3_energy
    h5md.dihedral_energy[frame] = bond4_energy
    h5md.field_energy[frame] = field_energy
    h5md.total_momentum[frame, :] = total_momentum
    h5md.angular_momentum[frame, :] = angular_momentum
    h5md.torque[frame, :] = torque
    h5md.temperature[frame] = temperature
    h5md.pressure[frame] = _host(pressure)
    box_size = _host(box_size).reshape(-1)
    for d in range(3):
        h5md.box_value[frame, d, d] = box_size[d]
    h5md.thermostat_work[frame] = thermostat_work

    # ---- the log line (file_io.py:705-777), column for column ----------------------------------------------------
    # (a fixed-width numpy string array on purpose: the reference builds its header in one, so labels longer than
    # "field E" are cut to 7 characters -- "fieldE/" in per-particle mode -- and tools parse what it prints)
    labels = np.array(["step", "time", "temp", "tot E", "kin E", "pot E", "field E", "elec E", "bond E", "ang E", "dih E",
                       "bias E", "Px", "Py", "Pz", "ΔH" if config.target_temperature else "ΔE"])
    show = np.ones(len(labels), dtype=bool)
    show[6:12] = np.array([field_energy, field_q_energy, bond2_energy, bond3_energy, bond4_energy, plumed_bias]) != 0.0
    n_cols = int(show.sum())
    initial_energy = getattr(config, "initial_energy", None)
    if initial_energy is None:
        labels[-1] = ""
    per = 1.0
    if dump_per_particle:
        for i in range(3, 9):
            labels[i] = labels[i][:-2] + "E/N"
        labels[-1] = labels[-1] + "/N"
        per = float(config.n_particles)
    if initial_energy is None:
        drift = 0.0
    elif config.target_temperature:
        drift = total_energy - initial_energy - thermostat_work
    else:
        drift = total_energy - initial_energy
    row = [step, time_step * step, temperature, total_energy / per, kinetic_energy / per, potential_energy / per,
           field_energy / per, field_q_energy / per, bond2_energy / per, bond3_energy / per, bond4_energy / per,
           plumed_bias / per, total_momentum[0] / per, total_momentum[1] / per, total_momentum[2] / per, drift / per]
    header = (n_cols * "{:>13}").format(*labels[show])
    data = ("{:13}" + (n_cols - 1) * "{:13.5g}").format(*[val for val, on in zip(row, show) if on])
    h5md.last_log = "\n" + header + "\n" + data
    if _world.current().rank == 0:
        log_rank0.info(h5md.last_log)


def distribute_input(in_file, rank, size, n_particles, max_molecule_size=201, comm=None):
    """Which rows of the input file rank ``rank`` of ``size`` reads (``file_io.py:780-874``): equal shares, moved to the
    next molecule boundary so that no molecule is split.  Returns ``(list of row indices, molecules present?)``."""
    if n_particles is None:
        n_particles = len(in_file["indices"])
    share = n_particles // size
    if "molecules" not in in_file:
        return list(range(rank * share, n_particles if rank == size - 1 else (rank + 1) * share)), False
    if size == 1:
        return list(range(0, n_particles)), True
    # look around the nominal break points for the places where the molecule id changes; like the reference this
    # assumes that no molecule is longer than min(max_molecule_size + 2, share) rows
    reach = min(max_molecule_size + 2, share)
    lo = 0 if rank == 0 else rank * share - 1
    hi = n_particles if rank == size - 1 else (rank + 1) * share + reach
    molecules = np.asarray(in_file["molecules"][lo:hi])
    indices = np.asarray(in_file["indices"][lo:hi])
    ends = np.nonzero(np.diff(molecules))[0]          # window row after which a new molecule starts
    if rank == 0:
        first = 0
    else:
        first = int(indices[ends[ends > 0][0]]) + 1
    if rank == size - 1:
        last = n_particles
    elif rank == 0:
        last = int(indices[ends[ends >= share][0] + 1])
    else:
        last = int(indices[ends[ends > share][0]]) + 1
    return list(range(first, last)), True


def read_input(path_or_file, config, dtype=np.float32, rank=None, size=None, topol=None):
    """The start-up reads of ``main.py:72-125``: this rank's rows of ``indices``, ``coordinates[-1]``,
    ``velocities[-1]`` (zeros if absent), ``names``, and when present ``types``, ``molecules``, ``bonds`` (only without
    a topology file), ``charge``; the ``box`` attribute overrides ``config.box_size``.  Returns a dict of numpy arrays
    (plus ``rank_range``, ``molecules_flag``); the caller moves what it wants to the device."""
    world = _world.current()
    rank = world.rank if rank is None else rank
    size = world.size if size is None else size
    opened = isinstance(path_or_file, (str, os.PathLike))
    in_file = _h5().File(path_or_file, "r") if opened else path_or_file
    try:
        rank_range, molecules_flag = distribute_input(
            in_file, rank, size, config.n_particles,
            config.max_molecule_size if getattr(config, "max_molecule_size", None) else 201)
        rows = slice(rank_range[0], rank_range[-1] + 1) if len(rank_range) else slice(0, 0)
        out = {"rank_range": rank_range, "molecules_flag": molecules_flag}
        out["indices"] = np.asarray(in_file["indices"][rows])
        out["positions"] = np.asarray(in_file["coordinates"][-1, rows, :]).astype(dtype)
        if "velocities" in in_file:
            out["velocities"] = np.asarray(in_file["velocities"][-1, rows, :]).astype(dtype)
        else:
            out["velocities"] = np.zeros_like(out["positions"], dtype=dtype)
        out["names"] = np.asarray(in_file["names"][rows])
        if "box" in in_file.attrs:
            config.box_size = np.array(in_file.attrs["box"])
        elif getattr(config, "box_size", None) is None:
            raise ValueError("No box size present in either config or input file. Unable to start simulation.")
        out["types"] = np.asarray(in_file["types"][rows]) if "types" in in_file else None
        out["molecules"] = np.asarray(in_file["molecules"][rows]) if molecules_flag else []
        out["bonds"] = np.asarray(in_file["bonds"][rows]) if molecules_flag and topol is None else None
        out["charges"] = np.asarray(in_file["charge"][rows]) if "charge" in in_file else None
        out["input_box"] = np.asarray(in_file["box"][:]) if "box" in in_file else np.array([None, None, None])
        return out
    finally:
        if opened:
            in_file.close()

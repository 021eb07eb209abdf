"""Synthetic systems for the benchmark configurations of BASELINE.json.

All systems: ``numpy.random.default_rng(seed)``, cubic box ``L = float32((N / 8.37)**(1/3))`` so
the number density is ~8.37 nm^-3 (``utils/units.py:5`` of the reference), kappa = 0.05,
``hamiltonian = "DefaultWithChi"``, ``m = [1.0] * T``, types assigned by fixed fractions.
Polymer types are laid out as 20-bead random-walk chains (bond 0.47 nm) wrapped periodically;
solvent is uniform.  Particle order follows HyMD's input format: the beads of a molecule are
consecutive (``distribute_input`` hands out whole molecules by index range,
``file_io.py:781-874``), molecules -- chains and single solvent beads alike -- appear in random
order, so the order carries no spatial information beyond the chains themselves.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from .config import Chi, Config

DENSITY = 8.37  # nm^-3

# DPPC lipid + water chi table (docs/doc_pages/examples.rst:449-460 of the reference)
DPPC_CHI = [("C", "W", 42.24), ("G", "C", 10.47), ("N", "W", -3.77), ("G", "W", 4.53),
            ("N", "P", -9.34), ("P", "G", 8.04), ("N", "G", 1.97), ("P", "C", 14.72),
            ("P", "W", -1.51), ("N", "C", 13.56)]


@dataclass
class System:
    name: str
    config: Config
    positions: np.ndarray      # (N,3) config.dtype
    types: np.ndarray          # (N,) int32
    charges: Optional[np.ndarray]
    velocities: np.ndarray


SPECS: Dict[str, dict] = {
    # ideal_chain-style 2-type homopolymer melt (examples.rst:371-373: chi_AB = 30)
    "C1": dict(n=10_000, mesh=24, names=["A", "B"], frac=[0.5, 0.5], polymer=[True, True],
               chi=[("A", "B", 30.0)], seed=1001),
    # DPPC bilayer + water: N,P,G,C per lipid = 1,1,2,8 plus water
    "C2": dict(n=100_000, mesh=64, names=["N", "P", "G", "C", "W"],
               frac=[1 / 38.5, 1 / 38.5, 2 / 38.5, 8 / 38.5, 26.5 / 38.5],
               polymer=[True, True, True, True, False], chi=DPPC_CHI, seed=1002),
    # charged lipid membrane + counter ions, PME
    "C3": dict(n=1_000_000, mesh=128, names=["N", "P", "G", "C", "W", "X"],
               frac=[1 / 39.5, 1 / 39.5, 2 / 39.5, 8 / 39.5, 26.5 / 39.5, 1 / 39.5],
               polymer=[True, True, True, True, False, False], chi=DPPC_CHI, seed=1003,
               charges={"N": 1.0, "P": -1.0, "X": 0.0}, coulomb=True),
    # multi-type polymer / water box
    "C4": dict(n=10_000_000, mesh=256, names=["A", "B", "C", "W"], frac=[0.2, 0.2, 0.1, 0.5],
               polymer=[True, True, True, False],
               chi=[("A", "B", 20.0), ("A", "C", -5.0), ("B", "C", 10.0), ("A", "W", 30.0),
                    ("B", "W", 5.0), ("C", "W", 0.0)], seed=1004),
    "C5": dict(n=100_000_000, mesh=512, names=["A", "B", "C", "W"], frac=[0.2, 0.2, 0.1, 0.5],
               polymer=[True, True, True, False],
               chi=[("A", "B", 20.0), ("A", "C", -5.0), ("B", "C", 10.0), ("A", "W", 30.0),
                    ("B", "W", 5.0), ("C", "W", 0.0)], seed=1005),
}


def box_length(n: int) -> float:
    return float(np.float32((n / DENSITY) ** (1.0 / 3.0)))


def make_system(name: str = "C2", dtype=np.float32, n: Optional[int] = None,
                mesh: Optional[int] = None, sigma: float = 0.5, kappa: float = 0.05,
                chains: bool = True, x_copies: int = 1, x_index: int = 0,
                x_chunk: Optional[Tuple[int, int, int]] = None) -> System:
    """Build one of the C1..C5 systems (optionally at a reduced ``n`` / ``mesh``).

    ``x_copies > 1`` (weak scaling): the global system is ``x_copies`` independent boxes of ``n``
    particles stacked along x (box ``[x_copies * L, L, L]``, mesh ``[x_copies * mesh, mesh,
    mesh]``); the call returns the particles of box ``x_index`` only (own seed) together with the
    GLOBAL config, so every rank generates just its own slab.

    ``x_chunk = (first, last, count)`` (strong scaling of the largest systems, where no rank can afford to
    generate all 1e8 particles): the cubic box is cut into ``count`` x-chunks, chunk ``k`` holds ``n / count``
    particles generated from its own seed with their chain starts / solvent positions uniform inside the
    chunk; the call returns chunks ``first .. last - 1``.  The GLOBAL system (all chunks) is the same however
    many ranks share it; chains may reach into the neighbouring chunk."""
    if x_chunk is not None:
        first, last, count = x_chunk
        parts = [_make_chunk(name, dtype, n, mesh, sigma, kappa, chains, k, count) for k in range(first, last)]
        return System(name=name, config=parts[0].config,
                      positions=np.concatenate([p.positions for p in parts]),
                      types=np.concatenate([p.types for p in parts]),
                      charges=None if parts[0].charges is None else np.concatenate([p.charges for p in parts]),
                      velocities=np.concatenate([p.velocities for p in parts]))
    spec = SPECS[name]
    n = int(n if n is not None else spec["n"])
    mesh = int(mesh if mesh is not None else spec["mesh"])
    rng = np.random.default_rng(spec["seed"] + 7919 * int(x_index))
    names: List[str] = spec["names"]
    L = box_length(n)
    frac = np.asarray(spec["frac"], dtype=np.float64)
    counts = np.floor(frac / frac.sum() * n).astype(np.int64)
    counts[-1 if not spec.get("coulomb") else len(names) - 2] += n - counts.sum()
    types = np.repeat(np.arange(len(names), dtype=np.int32), counts)
    pos = rng.uniform(0.0, L, size=(n, 3))
    if chains:
        # 20-bead random walks for the polymer types (vectorised: one cumulative sum per chain)
        for t, is_poly in enumerate(spec["polymer"]):
            if not is_poly:
                continue
            idx = np.nonzero(types == t)[0]
            nb = len(idx) // 20 * 20
            if nb == 0:
                continue
            steps = rng.normal(size=(nb // 20, 20, 3))
            steps *= 0.47 / np.linalg.norm(steps, axis=2, keepdims=True)
            steps[:, 0, :] = 0.0
            walk = np.cumsum(steps, axis=1) + pos[idx[:nb:20], None, :]
            pos[idx[:nb]] = np.mod(walk.reshape(nb, 3), L)
    # file order: molecules (20-bead chains, single solvent beads) in random order, the beads of
    # one molecule consecutive
    mol_start = []
    for t, is_poly in enumerate(spec["polymer"]):
        idx = np.nonzero(types == t)[0]
        if is_poly and chains:
            nb = len(idx) // 20 * 20
            mol_start.append(np.stack([idx[:nb:20], np.full(nb // 20, 20)], axis=1))
            rest = idx[nb:]
        else:
            rest = idx
        mol_start.append(np.stack([rest, np.ones(len(rest), dtype=np.int64)], axis=1))
    mols = np.concatenate(mol_start, axis=0)
    mols = mols[rng.permutation(len(mols))]
    ends = np.cumsum(mols[:, 1])
    begins = ends - mols[:, 1]
    perm = np.repeat(mols[:, 0] - begins, mols[:, 1]) + np.arange(n)
    pos, types = pos[perm], types[perm]
    charges = None
    coulomb = bool(spec.get("coulomb"))
    if coulomb:
        charges = np.zeros(n, dtype=np.float64)
        for nm, q in spec["charges"].items():
            charges[types == names.index(nm)] = q
        # counter ions neutralise exactly
        ions = np.nonzero(types == names.index("X"))[0]
        excess = charges.sum()
        k = min(len(ions), int(round(abs(excess))))
        charges[ions[:k]] = -np.sign(excess)
    vel = rng.normal(size=(n, 3)) * np.sqrt(Config.gas_constant * 323.0 / 72.0)
    cfg = Config(mesh_size=mesh if x_copies == 1 else [mesh * x_copies, mesh, mesh], sigma=sigma,
                 kappa=kappa, box_size=[L * x_copies, L, L],
                 hamiltonian="DefaultWithChi", chi=[Chi(*c) for c in spec["chi"]],
                 dtype=np.dtype(dtype), mass=72.0, time_step=0.01, respa_inner=25,
                 coulombtype="PIC_Spectral" if coulomb else None,
                 dielectric_const=80.0 if coulomb else None)
    cfg.finalize(names, n_particles=n * x_copies)
    if coulomb:
        cfg.type_charges = [spec["charges"].get(nm, 0.0) for nm in cfg.unique_names]
    # names sorted alphabetically define the type ids (input_parser.py:546-557): remap
    remap = np.array([cfg.name_to_type_map[nm] for nm in names], dtype=np.int32)
    types = remap[types]
    pos = np.mod(pos, L).astype(dtype)
    pos[pos >= np.asarray(L, dtype=dtype)] = 0.0
    if x_index:
        pos[:, 0] += np.asarray(L * x_index, dtype=dtype)
    return System(name=name, config=cfg, positions=pos, types=types.astype(np.int32),
                  charges=None if charges is None else charges.astype(dtype),
                  velocities=vel.astype(dtype))


def _make_chunk(name, dtype, n, mesh, sigma, kappa, chains, k, count):
    """Chunk ``k`` of ``count`` of the global system ``name`` (see ``make_system(x_chunk=...)``): the
    generator of the whole box run on ``n / count`` particles, squeezed into the chunk's x range."""
    spec = SPECS[name]
    n_glob = int(n if n is not None else spec["n"])
    sub = make_system(name, dtype=np.float64, n=n_glob // count, mesh=mesh, sigma=sigma, kappa=kappa, chains=False)
    L = box_length(n_glob)
    rng = np.random.default_rng(spec["seed"] + 104729 * (k + 1))
    m = len(sub.positions)
    pos = rng.uniform(0.0, 1.0, size=(m, 3)) * np.array([L / count, L, L]) + np.array([k * L / count, 0.0, 0.0])
    types = sub.types
    if chains:      # 20-bead random walks along the file order of every polymer type (vectorised)
        names = spec["names"]
        cfg0 = sub.config
        for t, is_poly in enumerate(spec["polymer"]):
            if not is_poly:
                continue
            idx = np.nonzero(types == cfg0.name_to_type_map[names[t]])[0]
            nb = len(idx) // 20 * 20
            if nb == 0:
                continue
            steps = rng.normal(size=(nb // 20, 20, 3))
            steps *= 0.47 / np.linalg.norm(steps, axis=2, keepdims=True)
            steps[:, 0, :] = 0.0
            walk = np.cumsum(steps, axis=1) + pos[idx[:nb:20], None, :]
            pos[idx[:nb]] = walk.reshape(nb, 3)
    pos = np.mod(pos, L).astype(dtype)
    pos[pos >= np.asarray(L, dtype=dtype)] = 0.0
    vel = rng.normal(size=(m, 3)) * np.sqrt(Config.gas_constant * 323.0 / 72.0)
    cfg = Config(mesh_size=int(mesh if mesh is not None else spec["mesh"]), sigma=sigma, kappa=kappa,
                 box_size=[L, L, L], hamiltonian="DefaultWithChi", chi=[Chi(*c) for c in spec["chi"]],
                 dtype=np.dtype(dtype), mass=72.0, time_step=0.01, respa_inner=25,
                 coulombtype=sub.config.coulombtype, dielectric_const=sub.config.dielectric_const)
    cfg.finalize(spec["names"], n_particles=(n_glob // count) * count)
    if sub.charges is not None:
        cfg.type_charges = sub.config.type_charges
    return System(name=name, config=cfg, positions=pos, types=types,
                  charges=None if sub.charges is None else sub.charges.astype(dtype), velocities=vel.astype(dtype))

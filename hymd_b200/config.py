"""Field-relevant slice of HyMD's ``Config`` (``hymd/input_parser.py:17-212``).

Only the keys the particle-mesh field-force cycle reads are modelled; names,
defaults and normalisation rules follow the reference so that a reference
``Config`` object and this one are interchangeable for ``hymd_b200.field``:

* ``mesh_size`` int or 3-list (``docs/doc_pages/config_file.rst:140-143``)
* ``box_size`` cast to float32 (``input_parser.py:736``)
* ``unique_names`` sorted, ``n_types`` (``input_parser.py:546-557``)
* ``type_to_name_map`` / ``name_to_type_map`` (``input_parser.py:560-582``)
* ``chi`` as ``Chi(atom_1, atom_2, interaction_energy)``; missing pairs are
  zero-filled (``input_parser.py:712-730``)
* ``m`` per-type paint mass, default 1.0 (``input_parser.py:1138-1141``)
* ``coulombtype``, ``dielectric_const``, ``type_charges``, ``self_energy``
* class constants ``coulomb_constant``, ``gas_constant``
  (``input_parser.py:159-160``)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import ClassVar, Dict, List, Optional, Sequence, Union

import numpy as np


@dataclass
class Chi:
    """``hymd/force.py:223-264``."""
    atom_1: str
    atom_2: str
    interaction_energy: float


@dataclass
class Config:
    gas_constant: ClassVar[float] = 0.0083144621  # kJ mol-1 K-1
    coulomb_constant: ClassVar[float] = 138.935458  # kJ nm mol-1 e-2

    n_steps: int = 1
    time_step: float = 0.01
    mesh_size: Union[int, Sequence[int], np.ndarray] = 32
    sigma: float = 0.5
    kappa: float = 0.05
    dtype: Optional[np.dtype] = None
    box_size: Optional[Union[Sequence[float], np.ndarray]] = None
    mass: float = 72.0
    hamiltonian: str = "DefaultNoChi"
    domain_decomposition: Union[int, bool, None] = None
    respa_inner: int = 1
    chi: List[Chi] = field(default_factory=list)
    n_particles: Optional[int] = None
    coulombtype: Optional[str] = None
    dielectric_const: Optional[float] = None
    self_energy: Optional[float] = None
    type_charges: Optional[Union[List[float], np.ndarray]] = None
    rho0: Optional[float] = None
    a: Optional[float] = None
    pressure: bool = False
    barostat: Optional[str] = None
    m: List[float] = field(default_factory=list)
    file_name: str = "<config file path unknown>"
    # set by finalize()
    unique_names: Optional[List[str]] = None
    n_types: Optional[int] = None
    type_to_name_map: Optional[Dict[int, str]] = None
    name_to_type_map: Optional[Dict[str, int]] = None

    def finalize(self, names: Sequence[str], n_particles: Optional[int] = None) -> "Config":
        """What ``check_config`` does for the field path (``input_parser.py:1296-1349``):
        sorted unique names -> type ids, zero-filled chi pairs, default ``m``,
        float32 box, default dtype."""
        names = [n.decode("utf-8") if isinstance(n, (bytes, np.bytes_)) else str(n) for n in names]
        self.unique_names = sorted(set(names))
        self.n_types = len(self.unique_names)
        self.name_to_type_map = {n: i for i, n in enumerate(self.unique_names)}
        self.type_to_name_map = {i: n for i, n in enumerate(self.unique_names)}
        if self.box_size is not None:
            self.box_size = np.array(self.box_size, dtype=np.float32)
        for i, n in enumerate(self.unique_names):
            for mname in self.unique_names[i + 1:]:
                if not any({c.atom_1, c.atom_2} == {n, mname} for c in self.chi):
                    self.chi.append(Chi(atom_1=n, atom_2=mname, interaction_energy=0.0))
        for c in self.chi:
            a, b = sorted([c.atom_1, c.atom_2])
            c.atom_1, c.atom_2 = a, b
        if not self.m:
            self.m = [1.0 for _ in range(self.n_types)]
        if n_particles is not None:
            self.n_particles = int(n_particles)
        if self.type_charges is None:
            self.type_charges = [0.0] * self.n_types
        if self.dtype is None:
            self.dtype = np.float32
        return self

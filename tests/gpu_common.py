"""Helpers shared by the GPU parity tests: run the CUDA path through the reference-shaped
Python API (which calls the C ABI) and the oracle on the same inputs."""
import numpy as np
import torch

from hymd_b200 import field as F
from hymd_b200.hamiltonian import get_hamiltonian
from oracle import field_oracle as fo
from oracle.hamiltonian_oracle import OracleHamiltonian


def rel_err(a, b):
    cplx = np.iscomplexobj(a) or np.iscomplexobj(b)
    a = np.asarray(a, dtype=np.complex128 if cplx else np.float64)
    b = np.asarray(b, dtype=np.complex128 if cplx else np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


class GpuRun:
    """One initialize_pm + update_field + compute_field_force (+ PME) pass on the GPU."""

    def __init__(self, cfg, positions, types, charges=None, compute_potential=False,
                 as_numpy=False):
        self.cfg = cfg
        self.h = get_hamiltonian(cfg)
        pm, fl, ecl, cl = F.initialize_pm(None, cfg)
        self.pm = pm
        (self.phi, self.phi_fourier, self.force_mesh, self.v_ext_fourier, self.v_ext,
         self.phi_transfer, self.phi_laplacian) = fl
        self.phi_q, self.phi_q_fourier, self.psi, self.elec_field = ecl
        tdt = torch.float64 if np.dtype(cfg.dtype) == np.float64 else torch.float32
        if as_numpy:
            self.pos, self.types = positions, types
            self.force = np.zeros((len(positions), 3), dtype=cfg.dtype)
        else:
            self.pos = torch.as_tensor(np.ascontiguousarray(positions), dtype=tdt, device="cuda")
            self.types = torch.as_tensor(types.astype(np.int32), device="cuda")
            self.force = torch.zeros((len(positions), 3), dtype=tdt, device="cuda")
        layouts = [pm.decompose(None) for _ in range(cfg.n_types)]
        F.update_field(self.phi, self.phi_laplacian, self.phi_transfer, layouts, self.force_mesh,
                       self.h, pm, self.pos, self.types, cfg, self.v_ext, self.phi_fourier,
                       self.v_ext_fourier, cfg.m, compute_potential=compute_potential)
        F.compute_field_force(layouts, self.pos, self.force_mesh, self.force, self.types,
                              cfg.n_types)
        self.elec_forces = None
        if charges is not None:
            if as_numpy:
                self.q = charges
                self.elec_forces = np.zeros((len(positions), 3), dtype=cfg.dtype)
            else:
                self.q = torch.as_tensor(charges, dtype=tdt, device="cuda")
                self.elec_forces = torch.zeros((len(positions), 3), dtype=tdt, device="cuda")
            F.update_field_force_q(self.q, self.phi_q, self.phi_q_fourier, self.psi, None, None,
                                   self.elec_field, self.elec_forces, pm.decompose(None), self.h,
                                   pm, self.pos, cfg)
        torch.cuda.synchronize()

    def forces(self):
        f = self.force
        return f.cpu().numpy() if isinstance(f, torch.Tensor) else f

    def eforces(self):
        f = self.elec_forces
        return f.cpu().numpy() if isinstance(f, torch.Tensor) else f

    def energies(self, velocities):
        return F.compute_field_and_kinetic_energy(
            self.phi, self.phi_q, self.psi, velocities, self.h, self.pos, self.types, self.v_ext,
            self.cfg, None)


class OracleRun:
    def __init__(self, cfg, positions, types, charges=None, dtype=np.float64,
                 compute_potential=True):
        import copy
        cfg = copy.deepcopy(cfg)
        self.cfg = cfg
        self.h = OracleHamiltonian(cfg)
        self.st = fo.FieldState(cfg, dtype)
        pos = np.asarray(positions, dtype=dtype)
        mt = len(pos) > 500_000      # full-size BASELINE configs: the threaded C loops (same arithmetic)
        fo.update_field(self.st, self.h, pos, types, cfg, compute_potential=compute_potential,
                        workers=-1, mt=mt)
        self.force = fo.compute_field_force(self.st, pos, types, cfg.n_types, mt=mt)
        self.elec_forces = None
        if charges is not None:
            self.elec_forces = fo.update_field_force_q(self.st, self.h, np.asarray(charges, dtype=dtype),
                                                       pos, cfg, workers=-1, mt=mt)

    def energies(self, velocities):
        return fo.compute_field_and_kinetic_energy(self.st, self.h, velocities, self.cfg)

"""Slab-sharded CPU restatement of the field-force cycle (x-slabs, one process per slab).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference distributes the mesh with pmesh/PFFT (``field.py:45-47``) and moves particles and
ghost contributions with ``pm.decompose`` / ``Layout.exchange`` (``main.py:977-980``,
``field.py:200, 574, 1165-1178``).  The B200 build replaces that by the data flow of
SURVEY.md section 8(e) / DESIGN.md section 4; this module restates THAT flow with numpy and
``torch.distributed`` (``gloo``) so that the decomposition logic can be checked on CPUs against
the single-rank oracle (``oracle/field_oracle.py``), which it must reproduce to round-off:

  1. every rank owns ``nxl = Nx / P`` x-planes of every real field and the particles whose cell
     lies in its slab;
  2. paint into ``nxl + 1`` planes (CIC touches plane i and i+1), **halo reduce**: the ghost
     plane is sent to rank+1 and added to its plane 0 (periodic);
  3. 2-D real FFT over (y, z) of the owned planes, **all-to-all** (ky blocks <-> x slabs),
     1-D FFT along x: the spectrum stays y-sharded ("k layout" ``[Nx][Ny/P][Nz/2+1]``);
  4. k-space arithmetic on the local block: filter, affine potential, ``-i k_d``;
  5. inverse: 1-D along x, all-to-all back, 2-D c2r over (y, z);
  6. **halo fetch**: plane 0 of rank+1 becomes this rank's ghost plane ``nxl``;
  7. readout of the local particles from ``nxl + 1`` planes, no x wrap needed.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as _fft

from . import pm_oracle as pmo


def affine_from_callables(hamiltonian, n_types, dtype=np.float64):
    """``V_t = sum_j A[t, j] phi_j + c[t]`` recovered by probing ``hamiltonian.v_ext``
    (``hamiltonian.py:402-412, 470-473``: all shipped functionals are affine)."""
    zero = [np.zeros(1, dtype=dtype) for _ in range(n_types)]
    c = np.array([float(np.asarray(hamiltonian.v_ext[t](zero)).reshape(-1)[0]) for t in range(n_types)])
    A = np.zeros((n_types, n_types))
    for j in range(n_types):
        probe = [np.full(1, 1.0 if i == j else 0.0, dtype=dtype) for i in range(n_types)]
        for t in range(n_types):
            A[t, j] = float(np.asarray(hamiltonian.v_ext[t](probe)).reshape(-1)[0]) - c[t]
    return A, c


class SlabComm:
    """The three exchanges of the sharded cycle over ``torch.distributed`` (any backend that has
    all_gather; gloo on CPUs)."""

    def __init__(self, dist):
        import torch
        self.dist, self.torch = dist, torch
        self.P, self.rank = dist.get_world_size(), dist.get_rank()

    def _gather(self, a):
        t = self.torch.from_numpy(np.ascontiguousarray(a))
        if np.iscomplexobj(a):
            t = self.torch.view_as_real(t)
        outs = [self.torch.empty_like(t) for _ in range(self.P)]
        self.dist.all_gather(outs, t)
        if np.iscomplexobj(a):
            outs = [self.torch.view_as_complex(o) for o in outs]
        return [o.numpy() for o in outs]

    def alltoall(self, blocks):
        """blocks[q] goes to rank q; returns the list of blocks received from ranks 0..P-1."""
        allb = self._gather(np.stack(blocks))
        return [allb[src][self.rank] for src in range(self.P)]

    def shift(self, a, direction):
        """Plane ``a`` travels to rank+direction (periodic); returns what arrives here."""
        allp = self._gather(a)
        return allp[(self.rank - direction) % self.P]

    def allreduce_sum(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return float(t.item())


def owner_of(positions, mesh, box, P):
    """Rank owning each particle: the slab of its CIC cell along x (what the GPU-side migration
    computes, ``hymd_migrate_plan``)."""
    nx = pmo.mesh_tuple(mesh)[0]
    x = np.asarray(positions, dtype=np.float64)[:, 0] * (nx / float(np.asarray(box)[0]))
    c = np.floor(x).astype(np.int64) % nx
    return c // (nx // P)


def _paint_slab(pos, mass, mesh, box, x0, nxl, dtype):
    """CIC into planes x0 .. x0+nxl (the last one is the ghost plane), periodic in y and z."""
    nx, ny, nz = mesh
    out = np.zeros((nxl + 1, ny, nz), dtype=dtype)
    if len(pos) == 0:
        return out
    n, c, d = pmo._cic_indices(pos, mesh, box, np.dtype(dtype))
    m = np.broadcast_to(np.asarray(mass, dtype=dtype), (len(pos),))
    one = np.dtype(dtype).type(1.0)
    lx = np.mod(c[:, 0], nx) - x0
    assert ((lx >= 0) & (lx < nxl)).all(), "particle outside its slab"
    flat = out.reshape(-1)
    for ax in (0, 1):
        for ay in (0, 1):
            for az in (0, 1):
                w = m * (d[:, 0] if ax else one - d[:, 0])
                w = w * (d[:, 1] if ay else one - d[:, 1])
                w = w * (d[:, 2] if az else one - d[:, 2])
                iy = np.mod(c[:, 1] + ay, ny)
                iz = np.mod(c[:, 2] + az, nz)
                np.add.at(flat, ((lx + ax) * ny + iy) * nz + iz, w.astype(dtype))
    return out


def _readout_slab(field_g, pos, mesh, box, x0):
    """CIC gather from ``nxl + 1`` planes (plane nxl = next slab's plane 0)."""
    nx, ny, nz = mesh
    dtype = field_g.dtype
    if len(pos) == 0:
        return np.zeros(0, dtype=dtype)
    n, c, d = pmo._cic_indices(pos, mesh, box, dtype)
    one = dtype.type(1.0)
    lx = np.mod(c[:, 0], nx) - x0
    out = np.zeros(len(pos), dtype=dtype)
    for ax in (0, 1):
        for ay in (0, 1):
            for az in (0, 1):
                w = (d[:, 0] if ax else one - d[:, 0])
                w = w * (d[:, 1] if ay else one - d[:, 1])
                w = w * (d[:, 2] if az else one - d[:, 2])
                out += w * field_g[lx + ax, np.mod(c[:, 1] + ay, ny), np.mod(c[:, 2] + az, nz)]
    return out


class SlabCycle:
    """One rank of the sharded cycle.  ``positions`` / ``types`` (/ ``charges``) are the LOCAL
    particles (every particle on the rank :func:`owner_of` names)."""

    def __init__(self, config, hamiltonian, comm: SlabComm, dtype=np.float64):
        self.cfg, self.h, self.comm = config, hamiltonian, comm
        self.dtype = np.dtype(dtype)
        self.mesh = pmo.mesh_tuple(config.mesh_size)
        self.box = np.asarray(config.box_size, dtype=np.float64)
        P = comm.P
        nx, ny, nz = self.mesh
        if nx % P or ny % P:
            raise ValueError("Nx and Ny must be divisible by the number of slabs")
        self.nxl, self.nyl = nx // P, ny // P
        self.x0, self.y0 = comm.rank * self.nxl, comm.rank * self.nyl
        kx, ky, kz = pmo.wavevectors(self.mesh, self.box, self.dtype)
        # k layout of this rank: all kx, ky block [y0, y0 + nyl), all stored kz
        self.k = [kx[:, None, None], ky[None, self.y0:self.y0 + self.nyl, None], kz[None, None, :]]
        self.M = nx * ny * nz

    # ---- distributed transforms ---------------------------------------------------------
    def forward(self, real_slab):
        """(nxl, Ny, Nz) real -> (Nx, nyl, Nzc) spectrum, 1/M normalised like ``r2c``."""
        P, nyl = self.comm.P, self.nyl
        a = _fft.rfft2(real_slab, axes=(1, 2))
        got = self.comm.alltoall([a[:, q * nyl:(q + 1) * nyl, :] for q in range(P)])
        k = np.concatenate(got, axis=0)                 # x slabs in rank order
        k = _fft.fft(k, axis=0)
        return (k * self.dtype.type(1.0 / self.M)).astype(np.result_type(self.dtype, np.complex64))

    def inverse(self, spec):
        """(Nx, nyl, Nzc) -> (nxl, Ny, Nz) real, unnormalised like ``c2r``."""
        P, nxl = self.comm.P, self.nxl
        nx, ny, nz = self.mesh
        a = _fft.ifft(spec, axis=0) * nx
        got = self.comm.alltoall([a[q * nxl:(q + 1) * nxl] for q in range(P)])
        planes = np.concatenate(got, axis=1)            # ky blocks in rank order
        out = _fft.irfft2(planes, s=(ny, nz), axes=(1, 2)) * (ny * nz)
        return out.astype(self.dtype)

    def with_ghost(self, slab):
        """Halo fetch: append plane 0 of rank+1."""
        ghost = self.comm.shift(slab[0], -1)
        return np.concatenate([slab, ghost[None]], axis=0)

    def paint(self, pos, mass):
        """Paint + halo reduce -> (nxl, Ny, Nz) owned planes."""
        g = _paint_slab(pos, mass, self.mesh, self.box, self.x0, self.nxl, self.dtype)
        incoming = self.comm.shift(g[self.nxl], +1)
        own = g[:self.nxl].copy()
        own[0] += incoming
        return own

    # ---- the cycle ----------------------------------------------------------------------
    def field_forces(self, positions, types):
        """``update_field`` + ``compute_field_force`` (``field.py:570-616, 197-200``) for the
        local particles; returns (forces (n_loc, 3), local filtered densities)."""
        cfg, dt = self.cfg, self.dtype
        T = cfg.n_types
        dv = float(np.prod(self.box) / self.M)
        m = cfg.m or [1.0] * T
        A, c = affine_from_callables(self.h, T)
        H = np.asarray(self.h.H(self.k, np.ones((1, 1, 1), dtype=dt)))     # filter on the local block
        phi_hat = []
        for t in range(T):
            rho = self.paint(positions[types == t], m[t]) / dt.type(dv)
            phi_hat.append(H * self.forward(rho))                           # field.py:576-577
        self.phi = [self.inverse(p) for p in phi_hat]                       # field.py:578
        origin = (self.y0 == 0)
        force = np.zeros((len(positions), 3), dtype=dt)
        for t in range(T):
            # v_ext is affine: its spectrum is the same combination plus c at k = 0
            vf = sum(A[t, j] * phi_hat[j] for j in range(T)).astype(phi_hat[0].dtype)
            if origin:
                vf[0, 0, 0] += c[t]
            vf = H * vf                                                     # field.py:585
            ind = types == t
            for d in range(3):
                g = self.inverse((-1j * self.k[d] * vf).astype(vf.dtype))   # field.py:607-613
                force[ind, d] = _readout_slab(self.with_ghost(g), positions[ind], self.mesh,
                                              self.box, self.x0)
        return force

    def pme_forces(self, positions, charges):
        """``update_field_force_q`` (``field.py:356-403``) for the local particles."""
        cfg, dt = self.cfg, self.dtype
        dv = float(np.prod(self.box) / self.M)
        conv = cfg.coulomb_constant / cfg.dielectric_const
        H = np.asarray(self.h.H(self.k, np.ones((1, 1, 1), dtype=dt)))
        q = np.asarray(charges, dtype=dt)
        rho = self.paint(positions, q) / dt.type(dv)
        pf = H * self.forward(rho)
        k2 = self.k[0] ** 2 + self.k[1] ** 2 + self.k[2] ** 2
        k2 = np.array(np.broadcast_to(k2, pf.shape), copy=True)
        if self.y0 == 0:
            k2[0, 0, 0] = 1.0                                               # normp(zeromode=1)
        out = np.zeros((len(positions), 3), dtype=dt)
        for d in range(3):
            ef = (-1j * self.k[d] * 4.0 * np.pi * conv * pf / k2).astype(pf.dtype)
            e = self.with_ghost(self.inverse(ef))
            out[:, d] = q * _readout_slab(e, positions, self.mesh, self.box, self.x0)
        self.phi_q = rho
        self.psi = self.inverse((4.0 * np.pi * conv * pf / k2).astype(pf.dtype))
        return out

    def field_energy(self):
        """``field.py:692-693`` summed over slabs."""
        dv = float(np.prod(self.box) / self.M)
        return self.comm.allreduce_sum(np.sum(self.h.w_0(self.phi) * dv, dtype=np.float64))

"""MD stepping around the field-force cycle (the caller side of the hot path).

``hymd/integrator.py:9-75`` (velocity Verlet) and the outer rRESPA loop of ``hymd/main.py:801-1169``:

* :class:`FieldOnlyMD` -- only the particle-field forces switched on (no bonds / angles / thermostat),
  on torch tensors or numpy arrays alike; drives the NVE energy-conservation check of the field-force
  path (``tools/nve_drift.py``, ``tests/test_gpu_nve.py``) with the CUDA path and with the oracle.
* :func:`kick_drift` and :class:`RespaMD` -- the full outer step with
  positions and velocities resident on the GPU: outer kicks by the slow (field, electrostatic) forces,
  ``respa_inner`` inner steps of bonded forces + kick + drift + wrap as fused kernels
  (``hymd_bonded_inner_step``, ``hymd_md_kick_drift``), the field-force cycle, and the thermostat
  (``hymd_b200.thermostat``).  Barostat, PLUMED and I/O stay with the caller.
"""
from __future__ import annotations


def integrate_velocity(velocities, accelerations, time_step):
    """Half kick ``v + dt/2 * a`` (``integrator.py:9-41``); returns a new array."""
    return velocities + (0.5 * time_step) * accelerations


def integrate_position(positions, velocities, time_step):
    """Drift ``x + dt * v`` (``integrator.py:44-75``); returns a new array."""
    return positions + time_step * velocities


def wrap(positions, box):
    """``np.mod(positions, box_size[None, :])`` (``main.py:837``) for numpy or torch.  A value
    that rounds up to exactly L (possible in floating point for tiny negative inputs) is mapped
    to 0 so positions stay in [0, L)."""
    try:
        import torch
        if isinstance(positions, torch.Tensor):
            b = torch.as_tensor(box, dtype=positions.dtype, device=positions.device)
            p = torch.remainder(positions, b)
            return torch.where(p >= b, torch.zeros_like(p), p)
    except ImportError:  # pragma: no cover
        pass
    import numpy as np
    b = np.asarray(box, dtype=positions.dtype)
    p = np.mod(positions, b[None, :])
    p[p >= b[None, :]] = 0
    return p


class FieldOnlyMD:
    """Outer-step propagator: ``force_fn(positions) -> field forces (N,3)`` is the field-force
    cycle (``update_field`` + ``compute_field_force``), everything else is the velocity-Verlet
    bookkeeping of ``main.py``."""

    def __init__(self, force_fn, box, mass, time_step, respa_inner=1):
        self.force_fn = force_fn
        self.box = box
        self.mass = float(mass)
        self.dt = float(time_step)
        self.inner = int(respa_inner)

    def step(self, positions, velocities, forces):
        """One outer step from (x, v, F(x)) to (x', v', F(x'))."""
        outer = self.inner * self.dt
        velocities = integrate_velocity(velocities, forces / self.mass, outer)   # main.py:803-807
        for _ in range(self.inner):                                               # main.py:829-837
            positions = wrap(integrate_position(positions, velocities, self.dt), self.box)
        forces = self.force_fn(positions)                                         # main.py:976-1004
        velocities = integrate_velocity(velocities, forces / self.mass, outer)   # main.py:1144-1148
        return positions, velocities, forces


# --- device-resident rRESPA step -----------------------------------

def kick_drift(vel, pos, forces, mass, kick_dt, drift_dt=0.0, box=None, sequential=False):
    """Fused velocity-Verlet update on CUDA tensors, in place (``csrc/md.cu``):
    ``vel += 0.5*kick_dt*(sum forces)/mass`` (or one kick per force array in turn with
    ``sequential``, the way ``main.py:803-827`` applies field / electrostatic / ... forces), then,
    if ``pos`` is given, ``pos = mod(pos + drift_dt*vel, box)`` (``main.py:835-837``) in the same
    pass over the arrays."""
    import ctypes

    import numpy as np
    import torch

    from . import _lib
    lib = _lib.load()
    forces = [f for f in forces if f is not None]
    for t in [vel] + forces + ([pos] if pos is not None else []):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype == vel.dtype
                and t.shape == vel.shape):
            raise ValueError("kick_drift needs contiguous CUDA tensors of one dtype and shape (N,3)")
    code = _lib.F64 if vel.dtype == torch.float64 else _lib.F32
    fptr = (ctypes.c_void_p * max(len(forces), 1))(*[f.data_ptr() for f in forces])
    bx = None
    if pos is not None:
        bx = (ctypes.c_double * 3)(*[float(b) for b in np.asarray(box).reshape(-1)[:3]])
    _lib.check(lib.hymd_md_kick_drift(
        code, ctypes.c_void_p(vel.data_ptr()), ctypes.c_void_p(pos.data_ptr()) if pos is not None else None,
        fptr, len(forces), 1 if sequential else 0, float(mass), float(kick_dt), float(drift_dt), bx,
        int(vel.shape[0]), ctypes.c_void_p(torch.cuda.current_stream(vel.device).cuda_stream)))
    _lib.mark_written(vel)
    if pos is not None:
        _lib.mark_written(pos)      # written through a raw pointer: invalidates cached bins (pm.sort)


class RespaMD:
    """One outer rRESPA step of ``main.py:764-1169`` with everything resident on the GPU.

    ``field_force_fn(positions) -> list of (N,3) slow-force tensors`` (field forces, and the
    electrostatic forces when charges are present: each gets its own kick, ``main.py:803-827``);
    ``topology`` is a :class:`hymd_b200.force.BondedTopology` (or ``None`` for monatomic systems).
    Positions and velocities are CUDA tensors updated in place; per inner step the launches are one
    fused kick+drift+wrap kernel, the bonded kernels and one kick kernel.  ``thermostat`` is a
    callable ``(velocities) -> None`` applied every ``n_b`` outer steps (``main.py:1290-1292``)."""

    def __init__(self, field_force_fn, box, mass, time_step, respa_inner=1, topology=None,
                 thermostat=None, n_b=1, fused=True, force_out=None, cta=1):
        self.field_force_fn = field_force_fn
        self.box = box
        self.mass = float(mass)
        self.dt = float(time_step)
        self.inner = int(respa_inner)
        self.topology = topology
        self.thermostat = thermostat
        self.n_b = int(n_b)
        self.step_count = 0
        self.fast = None
        self.bonded_results = {}
        self.fused = bool(fused)
        import os
        cta = int(os.environ.get("HYMD_B200_RESPA_CTA", cta))      # tuning override (tools/gpu_md_next.sh)
        self.cta = cta                  # term evaluation of the fused kernel: 1 = once per CTA (measured
                                        # fastest at C4, profiles/r1h_md_bench.json), 0 = per particle
        self.force_out = force_out      # optional [bond, angle, dihedral] (N,3) tensors filled at the
        self._x_alt = None              # end of every outer step by the fused path

    def fast_forces(self, positions):
        """Bond / angle / dihedral forces at ``positions`` (``main.py:841-887``): list of tensors."""
        import torch
        topo = self.topology
        if topo is None:
            return []
        if self.fast is None:
            self.fast = [torch.zeros_like(positions) if topo.n_terms[k] > 0 else None for k in range(3)]
        for k in range(3):
            if self.fast[k] is not None:
                self.bonded_results[k + 2] = topo.forces(k + 2, positions, self.box, self.fast[k])
        return [f for f in self.fast if f is not None]

    def bonded_energies(self):
        """{2: bond, 3: angle, 4: dihedral} energies of the last inner step (``main.py:964-970``);
        reads the device results (synchronizes)."""
        return {k: float(v[0].item()) for k, v in self.bonded_results.items()}

    def step(self, positions, velocities, slow_forces):
        """From ``(x, v, slow forces at x)`` to the next outer step, in place; returns the new slow
        forces.  With a topology the inner loop is ``respa_inner + 1`` launches of the fused kernel
        ``hymd_bonded_inner_step`` (forces + kick(s) + drift in one pass, positions double-buffered);
        ``fused=False`` (or no topology) runs the four separate launches per inner step."""
        outer = self.inner * self.dt
        kick_drift(velocities, None, slow_forces, self.mass, outer, sequential=True)     # main.py:803-827
        # (topologies with dihedrals of dih_type 1 run the per-particle variant of the fused kernel that carries
        # the bending term, whatever ``cta`` says)
        if self.topology is not None and self.fused:
            self._fused_inner(positions, velocities)
        else:
            if self.topology is not None and self.topology._cta not in (None, 0):
                self.topology.set_cta(0)     # the separate kernels are fastest per particle
            fast = self.fast_forces(positions) if self.fast is None else [f for f in self.fast if f is not None]
            for _ in range(self.inner):                                                  # main.py:829-893
                kick_drift(velocities, positions, fast, self.mass, self.dt, self.dt, self.box)
                fast = self.fast_forces(positions)
                kick_drift(velocities, None, fast, self.mass, self.dt)
        slow_forces = self.field_force_fn(positions)                                     # main.py:976-1058
        kick_drift(velocities, None, slow_forces, self.mass, outer, sequential=True)     # main.py:1144-1169
        # the reference tests np.mod(step, n_b) == 0 with step counting from 0 (main.py:1290-1292, 1315)
        if self.thermostat is not None and self.step_count % self.n_b == 0:
            self.thermostat(velocities)
        self.step_count += 1
        return slow_forces

    def _fused_inner(self, positions, velocities):
        """k d F k | k d F k | ... regrouped as [F k d] [F k k d] ... [F k]: the bonded forces at the
        current positions are evaluated inside the kernel that consumes them (the first launch
        re-evaluates the forces of the previous outer step's last positions instead of storing
        them), so no force array goes through HBM."""
        import torch
        topo = self.topology
        if self._x_alt is None or self._x_alt.shape != positions.shape or self._x_alt.dtype != positions.dtype:
            self._x_alt = torch.empty_like(positions)
        cur, alt = positions, self._x_alt
        for i in range(self.inner):
            topo.inner_step(cur, alt, velocities, self.box, self.mass, self.dt, 1 if i == 0 else 2, self.dt,
                            want_energies=False, cta=self.cta)
            cur, alt = alt, cur
        res = topo.inner_step(cur, None, velocities, self.box, self.mass, self.dt, 1, 0.0,
                              force_out=self.force_out, want_energies=True, cta=self.cta)
        self.bonded_results = {k + 2: res[k] for k in range(3) if topo.n_terms[k] > 0}
        if cur is not positions:
            positions.copy_(cur)

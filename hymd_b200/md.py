"""Field-only MD stepping around the field-force cycle (the caller side of the hot path).

``hymd/integrator.py:9-75`` (velocity Verlet) and the outer rRESPA skeleton of
``hymd/main.py:801-1148`` restated for device tensors, with only the particle-field forces
switched on (no bonds / angles / thermostat): kick by the field forces over the outer step
``respa_inner * time_step``, ``respa_inner`` drifts of ``time_step`` with periodic wrapping
(``main.py:836-837``), field-force cycle (``main.py:976-1004``), second kick
(``main.py:1144-1148``).  It exists to drive the NVE energy-conservation check of the
field-force path (``tools/nve_drift.py``, ``tests/test_gpu_nve.py``); bonded forces,
thermostat, barostat and I/O stay with the caller.

Every function works on torch tensors (GPU path) and on numpy arrays (oracle path) alike.
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

"""NVE energy conservation of the field-force path against the oracle (BASELINE.json
north_star: "NVE energy drift ... must match the reference's").  Short version of
tools/nve_drift.py (the 10k-step run is committed under profiles/): the same field-only velocity
Verlet trajectory (hymd_b200/md.py, main.py:801-1148 with respa_inner = 1) driven by the CUDA
path and by the CPU oracle.  fp64: the energy curves coincide (1e-9 relative); fp32: the drift
stays within 2e-5 of |E| per particle of the oracle's curve."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_nve_energy_curve_matches_oracle():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import nve_drift
    from hymd_b200.synthetic import make_system
    sysm = make_system("C1", dtype=np.float64, n=6000, mesh=16)
    n = len(sysm.positions)
    steps, every, dt = 60, 10, 0.01
    o, _ = nve_drift.run_oracle(sysm, steps, every, dt)
    g64, _ = nve_drift.run_gpu(sysm, np.float64, steps, every, dt)
    g32, _ = nve_drift.run_gpu(sysm, np.float32, steps, every, dt)
    eo = np.array([x[1] for x in o]) / n
    e64 = np.array([x[1] for x in g64]) / n
    e32 = np.array([x[1] for x in g32]) / n
    scale = np.abs(eo).max()
    assert np.abs(e64 - eo).max() / scale < 1e-9
    assert np.abs(e32 - eo).max() / scale < 2e-5
    # the integrator conserves energy to O(dt^2): the drift over the run is small against E_kin
    ekin = o[0][3] / n
    assert abs(eo[-1] - eo[0]) < 0.05 * ekin
    assert abs((e64[-1] - e64[0]) - (eo[-1] - eo[0])) / scale < 1e-9


@pytest.mark.timeout(900)
def test_nve_1000_steps_on_two_slabs_matches_oracle():
    """Row (g) with a longer, sharded gate: 1000 velocity-Verlet steps in fp64 on TWO slabs (virtual ranks
    on one GPU, tests/test_gpu_virtual_slabs.py) against the oracle.  The ranks own the particles by index
    parity and nothing re-homes them, so every step routes about half of the particles to the other slab and
    their forces back; the total energy curve (slab-summed field energy + kinetic energy) must coincide with
    the oracle's."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import nve_drift
    from hymd_b200.synthetic import make_system
    sysm = make_system("C1", dtype=np.float64, n=6000, mesh=16)
    n = len(sysm.positions)
    steps, every, dt = 1000, 100, 0.01
    o, _ = nve_drift.run_oracle(sysm, steps, every, dt)
    g, _ = nve_drift.run_gpu_slabs(sysm, np.float64, steps, every, dt, 2)
    eo = np.array([x[1] for x in o]) / n
    eg = np.array([x[1] for x in g]) / n
    scale = np.abs(eo).max()
    assert np.abs(eg - eo).max() / scale < 1e-9
    assert abs((eg[-1] - eg[0]) - (eo[-1] - eo[0])) / scale < 1e-9

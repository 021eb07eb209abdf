#!/usr/bin/env python
"""NVE energy drift of the field-force path (BASELINE.json north_star: "NVE energy drift over
10k steps must match the reference's").

    python tools/nve_drift.py [--workload C2] [--n 100000] [--mesh 64] [--steps 10000]
                              [--every 100] [--impl gpu-f32 gpu-f64 oracle] [--out file.json]

Field-only velocity Verlet (hymd_b200/md.py: the outer-step skeleton of main.py:801-1148 with
respa_inner = 1, time_step = 0.01 ps, no bonds, no thermostat) from the same initial state with

* gpu-f32 / gpu-f64 : the CUDA path through hymd_b200.field (positions stay on the GPU),
* oracle            : the CPU restatement of the reference path in fp64 (test infrastructure).

Every ``--every`` steps the total energy E = E_field + E_kin (field.py:688-703) is recorded;
the report holds, per implementation, drift(t) = (E(t) - E(0)) / N in kJ/mol per particle, its
end value, the RMS fluctuation and a linear-fit slope per 1000 steps, plus the GPU-vs-oracle
difference of E(t)/N.  The bar: the GPU drift matches the oracle's (same integrator, same
forces within 1e-5 / 1e-10), i.e. the fp64 curves coincide until round-off chaos separates the
trajectories and the fp32 curve stays within the fp64 curve's fluctuation band.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hymd_b200.md import FieldOnlyMD  # noqa: E402
from hymd_b200.synthetic import make_system  # noqa: E402


def run_gpu_slabs(sysm, dtype, steps, every, dt, P):
    """The same trajectory on P slabs (virtual ranks = threads of this process on one GPU): every rank
    owns an arbitrary 1/P share of the particles (index modulo P), nothing ever re-homes them, so every
    step routes the particles outside their owner's slab to the slab owner and their forces back."""
    from hymd_b200._world import VirtualRanks
    out = [None] * P

    def worker(rank):
        out[rank] = run_gpu(sysm, dtype, steps, every, dt, share=(rank, P))

    VirtualRanks(P).run(worker)
    return out[0]


def run_gpu(sysm, dtype, steps, every, dt, share=None):
    import torch
    from hymd_b200 import field as F
    from hymd_b200.hamiltonian import get_hamiltonian
    cfg = copy.deepcopy(sysm.config)
    cfg.dtype = np.dtype(dtype)
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    ham = get_hamiltonian(cfg)
    pm, fl, ecl, _ = F.initialize_pm(None, cfg)
    phi, phi_fourier, force_mesh, v_ext_fourier, v_ext, phi_transfer, phi_laplacian = fl
    layouts = [pm.decompose(None) for _ in range(cfg.n_types)]
    dev = pm.device
    sel = slice(None) if share is None else slice(share[0], None, share[1])
    pos = torch.as_tensor(np.ascontiguousarray(sysm.positions[sel].astype(dtype)), device=dev)
    vel = torch.as_tensor(np.ascontiguousarray(sysm.velocities[sel].astype(dtype)), device=dev)
    typ = torch.as_tensor(np.ascontiguousarray(sysm.types[sel].astype(np.int32)), device=dev)
    n = pos.shape[0]
    force = torch.zeros((n, 3), dtype=tdt, device=dev)

    def force_fn(p):
        p = p.contiguous()
        F.update_field(phi, phi_laplacian, phi_transfer, layouts, force_mesh, ham, pm, p, typ, cfg,
                       v_ext, phi_fourier, v_ext_fourier, cfg.m)
        F.compute_field_force(layouts, p, force_mesh, force, typ, cfg.n_types)
        force_fn.last = p
        return force.clone()

    def energy(p, v):
        e = F.compute_field_and_kinetic_energy(phi, None, None, v, ham, p, typ, v_ext, cfg, layouts)
        return e[0] + e[1], e[0], e[1]

    md = FieldOnlyMD(force_fn, cfg.box_size, cfg.mass, dt, 1)
    f = force_fn(pos)
    series = [(0,) + energy(pos, vel)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(1, steps + 1):
        pos, vel, f = md.step(pos, vel, f)
        if s % every == 0 or s == steps:
            series.append((s,) + energy(pos, vel))
    torch.cuda.synchronize()
    return series, time.perf_counter() - t0


def run_oracle(sysm, steps, every, dt):
    from oracle import field_oracle as fo
    from oracle.hamiltonian_oracle import OracleHamiltonian
    cfg = copy.deepcopy(sysm.config)
    cfg.dtype = np.dtype(np.float64)
    h = OracleHamiltonian(cfg)
    st = fo.FieldState(cfg, np.float64)
    pos = sysm.positions.astype(np.float64)
    vel = sysm.velocities.astype(np.float64)
    typ = sysm.types

    def force_fn(p):
        fo.update_field(st, h, p, typ, cfg, workers=-1, mt=True)
        return fo.compute_field_force(st, p, typ, cfg.n_types, mt=True)

    def energy(v):
        e = fo.compute_field_and_kinetic_energy(st, h, v, cfg)
        return float(e[0] + e[1]), float(e[0]), float(e[1])

    md = FieldOnlyMD(force_fn, cfg.box_size, cfg.mass, dt, 1)
    f = force_fn(pos)
    series = [(0,) + energy(vel)]
    t0 = time.perf_counter()
    for s in range(1, steps + 1):
        pos, vel, f = md.step(pos, vel, f)
        if s % every == 0 or s == steps:
            series.append((s,) + energy(vel))
    return series, time.perf_counter() - t0


def summarize(series, n):
    s = np.asarray([x[0] for x in series], dtype=np.float64)
    e = np.asarray([x[1] for x in series], dtype=np.float64)
    drift = (e - e[0]) / n
    slope = float(np.polyfit(s, drift, 1)[0] * 1000.0) if len(s) > 2 else 0.0
    return {"steps": [int(x) for x in s], "energy_per_particle": (e / n).tolist(),
            "drift_per_particle": drift.tolist(), "drift_end": float(drift[-1]),
            "drift_rms": float(np.sqrt(np.mean((drift - drift.mean()) ** 2))),
            "drift_slope_per_1000_steps": slope,
            "e_field0_per_particle": series[0][2] / n, "e_kin0_per_particle": series[0][3] / n}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--mesh", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10_000)
    ap.add_argument("--every", type=int, default=100)
    ap.add_argument("--dt", type=float, default=0.01)
    ap.add_argument("--impl", nargs="+", default=["gpu-f32", "gpu-f64", "oracle"])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    sysm = make_system(a.workload, dtype=np.float64, n=a.n, mesh=a.mesh)
    n = len(sysm.positions)
    rep = {"workload": f"{a.workload} reduced to N={n}, mesh {a.mesh}^3, field forces only, NVE, "
                       f"velocity Verlet dt={a.dt} ps, respa_inner=1, {a.steps} steps",
           "unit": "kJ/mol per particle", "runs": {}}
    for impl in a.impl:
        if impl == "oracle":
            series, wall = run_oracle(sysm, a.steps, a.every, a.dt)
        else:
            series, wall = run_gpu(sysm, np.float32 if impl.endswith("f32") else np.float64, a.steps,
                                   a.every, a.dt)
        r = summarize(series, n)
        r["wall_s"] = wall
        rep["runs"][impl] = r
        print(f"{impl:8s} {a.steps} steps in {wall:7.1f} s  drift_end {r['drift_end']:+.3e}  "
              f"rms {r['drift_rms']:.3e}  slope/1000 {r['drift_slope_per_1000_steps']:+.3e}", flush=True)
    if "oracle" in rep["runs"]:
        eo = np.asarray(rep["runs"]["oracle"]["energy_per_particle"])
        for impl, r in rep["runs"].items():
            if impl != "oracle":
                d = np.abs(np.asarray(r["energy_per_particle"]) - eo)
                r["max_abs_diff_vs_oracle"] = float(d.max())
                r["first_diff_vs_oracle"] = float(d[min(1, len(d) - 1)])
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as fh:
            json.dump(rep, fh, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())

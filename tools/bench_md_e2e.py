"""End-to-end outer MD step on one GPU: field-force cycle + respa_inner fused inner steps + CSVR.

    python tools/bench_md_e2e.py [--n 10000000] [--mesh 256] [--steps 3] [--inner 25] [--out file.json]

The C4 system of SURVEY.md section 8d with its bonded side switched on: half of the particles in 20-bead
chains of types A / B / C (0.2 / 0.2 / 0.1 of all particles; bonds 0.47 nm / 1250, straight angles / 25),
half solvent W, chi table of C4, 256^3 mesh, time_step 0.01 ps, respa_inner 25 (the lipid example's
values, examples.rst:426-431), CSVR thermostat every outer step with two coupling groups.  Particles are
laid out as domain_decomposition returns them (molecules contiguous, ordered by the mesh cell of their
first bead).  Everything stays on the device (hymd_b200.md.RespaMD); the timed region is `steps` outer
steps between CUDA events after one warm-up step.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hymd_b200 import field as F  # noqa: E402
from hymd_b200 import thermostat as T  # noqa: E402
from hymd_b200.config import Chi, Config  # noqa: E402
from hymd_b200.force import BondedTopology  # noqa: E402
from hymd_b200.hamiltonian import get_hamiltonian  # noqa: E402
from hymd_b200.md import RespaMD  # noqa: E402
from hymd_b200.synthetic import SPECS  # noqa: E402


def build_system(n, mesh, rng, order="cell", block=16):
    L = float(np.float32((n / 8.37) ** (1.0 / 3.0)))
    nch = n // 40
    steps = rng.normal(size=(nch, 20, 3)).astype(np.float32)
    steps *= 0.47 / np.linalg.norm(steps, axis=2, keepdims=True)
    steps[:, 0] = rng.uniform(0, L, size=(nch, 3))
    chains = np.mod(np.cumsum(steps, axis=1), L).astype(np.float32)
    solvent = rng.uniform(0, L, size=(n - nch * 20, 3)).astype(np.float32)
    ctype = np.array([0, 0, 1, 1, 2], dtype=np.int32)[np.arange(nch) % 5]          # A A B B C
    # molecules in the cell order of their first bead (what domain_decomposition hands back)
    first = np.concatenate([chains[:, 0], solvent])
    cell = np.minimum((first * (mesh / L)).astype(np.int64), mesh - 1)
    key = (cell[:, 0] * mesh + cell[:, 1]) * mesh + cell[:, 2]
    if order == "block":
        # molecules ordered by (block of block^3 cells, chains before solvent, cell): CTAs of the bonded
        # kernels become homogeneous again while the field kernels keep block-level locality
        nb = (mesh + block - 1) // block
        blk = ((cell[:, 0] // block) * nb + cell[:, 1] // block) * nb + cell[:, 2] // block
        is_solvent = (np.arange(len(first)) >= nch).astype(np.int64)
        key = (blk * 2 + is_solvent) * (mesh ** 3) + key
    order = np.argsort(key, kind="stable")
    mlen = np.concatenate([np.full(nch, 20, dtype=np.int64), np.ones(len(solvent), dtype=np.int64)])[order]
    start = np.cumsum(mlen) - mlen
    pos = np.empty((n, 3), dtype=np.float32)
    types = np.empty(n, dtype=np.int32)
    is_chain = order < nch
    cs = start[is_chain]
    idx = (cs[:, None] + np.arange(20)[None, :]).ravel()
    pos[idx] = chains[order[is_chain]].reshape(-1, 3)
    types[idx] = np.repeat(ctype[order[is_chain]], 20)
    ss = start[~is_chain]
    pos[ss] = solvent[order[~is_chain] - nch]
    types[ss] = 3
    pos[pos >= np.float32(L)] = 0.0
    a2 = (cs[:, None] + np.arange(19)[None, :]).ravel().astype(np.int32)
    a3 = (cs[:, None] + np.arange(18)[None, :]).ravel().astype(np.int32)
    return L, pos, types, a2, a3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--mesh", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--inner", type=int, default=25)
    ap.add_argument("--order", choices=["cell", "block"], default="cell",
                    help="particle order: cell = what domain_decomposition returns today; block = molecules "
                         "grouped by (16^3-cell block of ~2400 particles, has bonds) -- DESIGN.md section 8")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    t0 = time.time()
    rng = np.random.default_rng(1004)
    L, pos, types, a2, a3 = build_system(args.n, args.mesh, rng, args.order)
    n = args.n
    names = ["A", "B", "C", "W"]
    cfg = Config(mesh_size=args.mesh, sigma=0.5, kappa=0.05, box_size=[L, L, L], hamiltonian="DefaultWithChi",
                 chi=[Chi(*c) for c in SPECS["C4"]["chi"]], dtype=np.dtype(np.float32), mass=72.0,
                 time_step=0.01, respa_inner=args.inner)
    cfg.finalize(names, n_particles=n)
    cfg.target_temperature, cfg.tau, cfg.thermostat_work = 323.0, 0.1, 0.0
    cfg.thermostat_coupling_groups = [["A", "B", "C"], ["W"]]
    ham = get_hamiltonian(cfg)
    pm, fl, _, _ = F.initialize_pm(None, cfg)
    phi, phi_fourier, force_mesh, v_ext_fourier, v_ext, phi_transfer, phi_laplacian = fl
    layouts = [pm.decompose(None) for _ in range(cfg.n_types)]
    dev = pm.device
    x = torch.as_tensor(pos, device=dev)
    typ = torch.as_tensor(types, device=dev)
    v = torch.randn((n, 3), device=dev) * float(np.sqrt(Config.gas_constant * 323.0 / 72.0))
    force = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    topo = BondedTopology(n, bonds=(a2, a2 + 1, np.full(len(a2), 0.47), np.full(len(a2), 1250.0)),
                          angles=(a3, a3 + 1, a3 + 2, np.full(len(a3), np.pi), np.full(len(a3), 25.0)),
                          device=dev)
    prng = np.random.default_rng(7)

    def field_forces(p):
        F.update_field(phi, phi_laplacian, phi_transfer, layouts, force_mesh, ham, pm, p, typ, cfg, v_ext,
                       phi_fourier, v_ext_fourier, cfg.m)
        F.compute_field_force(layouts, p, force_mesh, force, typ, cfg.n_types)
        return [force]

    md = RespaMD(field_forces, cfg.box_size, cfg.mass, cfg.time_step, respa_inner=args.inner, topology=topo,
                 thermostat=lambda vel: T.csvr_thermostat(vel, typ, cfg, prng), n_b=1)
    slow = field_forces(x)
    slow = md.step(x, v, slow)                       # warm-up
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    l0 = pm.launch_count() + topo.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        slow = md.step(x, v, slow)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    ps_per_step = cfg.time_step * args.inner
    en = md.bonded_energies()
    temp = float(T.kinetic_energy(v, cfg.mass)) * 2.0 / (3.0 * Config.gas_constant * n)
    res = {"workload": f"C4 + bonded chains: N={n}, mesh {args.mesh}^3, T=4, respa_inner={args.inner}, CSVR",
           "order": args.order,
           "ms_per_outer_step": ms, "outer_steps_per_s": 1e3 / ms,
           "ns_per_day": 1e3 / ms * ps_per_step * 1e-3 * 86400.0,
           "particle_steps_per_s": n * 1e3 / ms, "bonds": int(len(a2)), "angles": int(len(a3)),
           "launches_per_outer_step": (pm.launch_count() + topo.launch_count() - l0) / args.steps,
           "bond_energy": en.get(2), "angle_energy": en.get(3), "temperature_K": temp,
           "finite": bool(torch.isfinite(x).all() and torch.isfinite(v).all()), "setup_s": setup_s}
    print(json.dumps(res))
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()

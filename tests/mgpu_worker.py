"""Worker for the slab-sharded parity tests (run under torchrun, one rank per GPU; also runs with
a single process to exercise the slab pipeline on one GPU via HYMD_B200_FORCE_SLAB=1).

    torchrun --nproc-per-node P tests/mgpu_worker.py [--dtype f64] [--pme] [--mesh 32 24 40]
                                                     [--particles 20000] [--migrate] [--route]

Every rank builds the same seeded system, keeps the particles of its x-slab (or, with
--migrate, an arbitrary 1/P share that is re-homed by domain_decomposition; with --route the
same share WITHOUT re-homing: most particles are then guests on another rank's slab and are routed
there and back inside every field call), runs update_field + compute_field_force (+ PME) through
hymd_b200.field, and rank 0 compares the gathered forces, filtered densities and force meshes
with the CPU oracle at the north-star tolerances (1e-5 fp32, 1e-10 fp64)."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import make_config  # noqa: E402
from gpu_common import OracleRun, rel_err  # noqa: E402
from hymd_b200 import field as F  # noqa: E402
from hymd_b200.hamiltonian import get_hamiltonian  # noqa: E402

CHI3 = [("A", "B", 9.6754032616815161), ("A", "C", -13.2596290315913623),
        ("B", "C", 0.3852001771213374)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--pme", action="store_true")
    ap.add_argument("--mesh", type=int, nargs=3, default=[32, 24, 40])
    ap.add_argument("--particles", dest="n", type=int, default=20000)
    ap.add_argument("--migrate", action="store_true")
    ap.add_argument("--route", action="store_true",
                    help="arbitrary share WITHOUT re-homing: the per-step routing inside the library must "
                         "make the cycle equal the oracle")
    ap.add_argument("--seed", type=int, default=11)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype = np.float64 if args.dtype == "f64" else np.float32
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    tol = 1e-10 if args.dtype == "f64" else 1e-5

    rng = np.random.default_rng(args.seed)
    box = np.asarray([4.0, 5.0, 6.0], dtype=np.float32)
    n = args.n
    pos = (rng.uniform(0, 1, size=(n, 3)) * box).astype(dtype)
    pos = np.minimum(pos, np.nextafter(box.astype(dtype), 0).astype(dtype))
    names = ["ABC"[i % 3] for i in range(n)]
    cfg = make_config(names, n, args.mesh, box, chi=CHI3, dtype=dtype,
                      coulombtype="PIC_Spectral" if args.pme else None,
                      dielectric_const=80.0 if args.pme else None)
    types = np.array([cfg.name_to_type_map[t] for t in names], dtype=np.int32)
    q = None
    if args.pme:
        q = rng.choice([-1.0, 0.0, 1.0], size=n)
        q -= q.mean()
        q = q.astype(dtype)

    # ownership
    nxl = args.mesh[0] // world
    cell = np.floor(pos[:, 0].astype(np.float64) * args.mesh[0] / float(box[0])).astype(np.int64) % args.mesh[0]
    if args.migrate or args.route:
        mine = (np.arange(n) % world) == rank          # arbitrary share, most particles not home
    else:
        mine = (cell // nxl) == rank
    idx = np.nonzero(mine)[0]

    ham = get_hamiltonian(cfg)
    pm, fl, ecl, cl = F.initialize_pm(None, cfg)
    phi, phi_fourier, force_mesh, v_ext_fourier, v_ext, phi_transfer, phi_laplacian = fl
    phi_q, phi_q_fourier, psi, elec_field = ecl
    dev = pm.device
    pos_d = torch.as_tensor(np.ascontiguousarray(pos[idx]), dtype=tdt, device=dev)
    typ_d = torch.as_tensor(types[idx], device=dev)
    gid_d = torch.as_tensor(idx.astype(np.int64), device=dev)
    q_d = None if q is None else torch.as_tensor(q[idx], dtype=tdt, device=dev)
    if args.migrate:
        extra = (typ_d, gid_d) if q_d is None else (typ_d, gid_d, q_d)
        out = F.domain_decomposition(pos_d, pm, *extra)
        pos_d, typ_d, gid_d = out[0], out[1], out[2]
        if q_d is not None:
            q_d = out[3]
        st_cells = torch.floor(pos_d[:, 0].double() * args.mesh[0] / float(box[0])).long() % args.mesh[0]
        assert bool(((st_cells // nxl) == rank).all()), "migrate left particles outside the slab"
    layouts = [pm.decompose(None) for _ in range(cfg.n_types)]
    force_d = torch.zeros((len(pos_d), 3), dtype=tdt, device=dev)
    F.update_field(phi, phi_laplacian, phi_transfer, layouts, force_mesh, ham, pm, pos_d, typ_d,
                   cfg, v_ext, phi_fourier, v_ext_fourier, cfg.m, compute_potential=True)
    F.compute_field_force(layouts, pos_d, force_mesh, force_d, typ_d, cfg.n_types)
    eforce_d = None
    if args.pme:
        eforce_d = torch.zeros((len(pos_d), 3), dtype=tdt, device=dev)
        F.update_field_force_q(q_d, phi_q, phi_q_fourier, psi, None, None, elec_field, eforce_d,
                               pm.decompose(None), ham, pm, pos_d, cfg)
    vel = torch.zeros_like(pos_d)
    energies = F.compute_field_and_kinetic_energy(phi, phi_q, psi, vel, ham, pos_d, typ_d, v_ext,
                                                  cfg, layouts)
    # by-products of the print step: Laplacians and the pressure vector (all ranks, collective)
    from hymd_b200.pressure import comp_pressure
    bond_pr, angle_pr = np.array([1.0, -2.0, 0.5]) / world, np.array([0.25, 0.5, -0.75]) / world
    pressure = comp_pressure(phi, phi_q, psi if args.pme else None, ham, vel, cfg, phi_fourier,
                             phi_laplacian, phi_transfer, pos_d, bond_pr, angle_pr)
    torch.cuda.synchronize()

    payload = {
        "gid": gid_d.cpu().numpy(), "force": force_d.cpu().numpy(),
        "eforce": None if eforce_d is None else eforce_d.cpu().numpy(),
        "phi": [p.value.cpu().numpy() for p in phi],
        "fmesh": [[force_mesh[t][d].value.cpu().numpy() for d in range(3)] for t in range(cfg.n_types)],
        "v_ext": [v.value.cpu().numpy() for v in v_ext],
        "psi": None if not args.pme else psi.value.cpu().numpy(),
        "phi_fourier": [p.value.cpu().numpy() for p in phi_fourier],
        "lap": [[phi_laplacian[t][d].value.cpu().numpy() for d in range(3)] for t in range(cfg.n_types)],
    }
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, payload)
    else:
        gathered = [payload]
    ok = True
    if rank == 0:
        o = OracleRun(cfg, pos, types, charges=q)
        force = np.zeros((n, 3))
        eforce = np.zeros((n, 3))
        seen = np.zeros(n, dtype=np.int64)
        for g in gathered:
            force[g["gid"]] = g["force"]
            seen[g["gid"]] += 1
            if args.pme:
                eforce[g["gid"]] = g["eforce"]
        assert (seen == 1).all(), "particles lost or duplicated"
        checks = {"force": rel_err(force, o.force)}
        if args.pme:
            checks["eforce"] = rel_err(eforce, o.elec_forces)
            checks["psi"] = rel_err(np.concatenate([g["psi"] for g in gathered], axis=0), o.st.psi)
        for t in range(cfg.n_types):
            checks[f"phi{t}"] = rel_err(np.concatenate([g["phi"][t] for g in gathered], axis=0), o.st.phi[t])
            scale = np.abs(o.st.v_ext[t]).max() + 1.0 / cfg.kappa
            checks[f"v_ext{t}"] = np.abs(np.concatenate([g["v_ext"][t] for g in gathered], axis=0)
                                         - o.st.v_ext[t]).max() / scale
            # k-space is sharded along y in the slab layout
            ax = 1 if world > 1 else 0
            checks[f"phi_fourier{t}"] = rel_err(
                np.concatenate([g["phi_fourier"][t] for g in gathered], axis=ax), o.st.phi_fourier[t])
            for d in range(3):
                checks[f"fmesh{t}{d}"] = rel_err(
                    np.concatenate([g["fmesh"][t][d] for g in gathered], axis=0), o.st.force_mesh[t][d])
        e_o = o.energies(np.zeros_like(pos))
        escale = max(abs(e_o[0]), 0.5 * n / cfg.kappa * 1e-2)
        checks["field_energy"] = abs(energies[0] - e_o[0]) / escale
        if args.pme:
            checks["field_q_energy"] = abs(energies[2] - e_o[2]) / max(abs(e_o[2]), 1e-300)
        from oracle import field_oracle as fo
        want_p = fo.comp_pressure(o.st, o.h, np.zeros_like(pos), o.cfg, bond_pr * world, angle_pr * world)
        checks["pressure"] = np.abs(pressure - want_p).max() / np.abs(want_p).max() / 10.0
        lscale = max(np.abs(o.st.phi_laplacian[t][d]).max() for t in range(cfg.n_types) for d in range(3))
        for t in range(cfg.n_types):
            for d in range(3):
                got = np.concatenate([g["lap"][t][d] for g in gathered], axis=0)
                checks[f"lap{t}{d}"] = np.abs(got - o.st.phi_laplacian[t][d]).max() / lscale
        worst = max(checks.values())
        ok = worst < tol
        print(f"MGPU world={world} dtype={args.dtype} pme={args.pme} mesh={args.mesh} "
              f"worst={worst:.3e} tol={tol:g} {'OK' if ok else 'FAIL'}", flush=True)
        if not ok:
            for k, v in checks.items():
                if v >= tol:
                    print("  ", k, v, flush=True)
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, src=0)
        ok = bool(flag.item())
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the particle-mesh field-force cycle (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload C4] [--dtype f32|f64] [--scaling strong|weak]

One "step" = one field-force cycle on one batch of synthetic particles: cell binning
(pm.decompose) + update_field (paint, FFTs, fused k-space) + compute_field_force (readout)
[+ update_field_force_q for the PME workload C3], i.e. what HyMD's main.py:976-1058 runs every
outer MD step.  Prints ONE JSON line (rank 0).

* value      particle-steps/s with positions/types already resident in HBM (CUDA events,
             max over ranks), through hymd_b200.field (the reference-shaped API over the C ABI)
* e2e        same metric with HOST (pinned) positions/types in and HOST forces out every step
* roofline   dominant hand-written kernel: algorithmic bytes / CUDA-event time vs measured HBM
* phases     every phase of the cycle with its algorithmic bytes, time and GB/s
* cpu_baseline  the CPU restatement of the reference path (oracle/, "port") timed on this
             box's host cores on a bounded sample (N = 1 only)

--impl reference times that CPU port instead (the reference's own pmesh/PFFT/mpi4py stack
cannot be installed in this image), on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "field-force particle-steps/s"
UNIT = "particle-steps/s"
PS_PER_CYCLE = 0.01 * 25        # time_step 0.01 ps x respa_inner 25 (examples.rst:426-431)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--n", type=int, default=None, help="override particle count (debug)")
    ap.add_argument("--mesh", type=int, default=None, help="override mesh size (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--whole", action="store_true",
                    help="C5: every rank generates the whole 1e8-particle box (default: x-chunks with their own "
                         "seeds, each rank generates only its share)")
    ap.add_argument("--parity-max-n", type=int, default=25_000_000,
                    help="largest global particle count for which rank 0 runs a CPU oracle cycle "
                         "and compares every particle's force (above: momentum conservation only)")
    ap.add_argument("--no-dd", action="store_true",
                    help="skip the start-up domain_decomposition call (main.py:274-300): the "
                         "particles then stay in the generator's molecule order instead of the "
                         "mesh-cell order domain_decomposition hands back")
    return ap.parse_args()


def peaks():
    """Measured roofline denominators (MEASURED_PEAKS.json), else the profiling guide's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, dtype):
    """DRAM bytes per launch of each hand-written kernel from the committed ncu --set full
    capture (profiles/traffic.json), or {} when there is none for this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)
        return t.get(f"{workload}-{dtype}", {})
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 "200", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def build_system(args, world, rank=0):
    """Strong scaling: the named workload, every rank keeps the particles of its x-slab.
    Weak scaling: `world` independent boxes of the named workload stacked along x (mesh
    [world * M, M, M], N * world particles: per-GPU work fixed, power-of-two planes kept);
    every rank generates its own box only."""
    from hymd_b200.synthetic import make_system
    dtype = np.float32 if args.dtype == "f32" else np.float64
    if args.scaling == "weak" and world > 1:
        return make_system(args.workload, dtype=dtype, n=args.n, mesh=args.mesh, x_copies=world,
                           x_index=rank), True
    if args.workload == "C5" and not args.whole:
        # 1e8 particles: no rank generates the whole box.  The global system is 8 x-chunks with their own
        # seeds (hymd_b200.synthetic.make_system, x_chunk); rank r generates chunks [8 r / P, 8 (r+1) / P),
        # i.e. the same global system for P = 1, 2, 4, 8.
        per = CHUNKS // world
        return make_system(args.workload, dtype=dtype, n=args.n, mesh=args.mesh,
                           x_chunk=(rank * per, (rank + 1) * per, CHUNKS)), True
    return make_system(args.workload, dtype=dtype, n=args.n, mesh=args.mesh), False


def algorithmic_bytes(N, T, U, mesh, b, pme):
    """bytes each phase must move once (force-only cycle)."""
    M = int(np.prod(mesh))
    Mc = mesh[0] * mesh[1] * (mesh[2] // 2 + 1)
    out = {
        "paint": N * (3 * b + 4) + T * M * b,
        "fft_fwd": T * (M * b + Mc * 2 * b),
        "kspace": T * Mc * 2 * b + 3 * U * Mc * 2 * b,
        "fft_inv": 3 * U * (Mc * 2 * b + M * b),
        "readout": N * (3 * b + 4) + 3 * U * M * b + N * 3 * b,
    }
    if pme:
        out["pme_paint"] = N * (4 * b + 4) + M * b
        out["pme_kspace"] = Mc * 2 * b + 3 * Mc * 2 * b
        out["pme_fft"] = (M * b + Mc * 2 * b) + 3 * (Mc * 2 * b + M * b)
        out["pme_readout"] = N * (4 * b + 4) + 3 * M * b + N * 3 * b
    return out


def workload_config(args, world, cfg, n_global, pme):
    """The `config` object of the JSON line: identical for the GPU arm and the reference arm of
    one invocation (the driver compares them)."""
    from hymd_b200.hamiltonian import affine_parameters, get_hamiltonian
    mesh = [int(x) for x in np.full(3, cfg.mesh_size)]
    T = cfg.n_types
    A, c = affine_parameters(get_hamiltonian(cfg), T)
    U = len({(tuple(A[t]), c[t]) for t in range(T)})
    return {"workload": f"{args.workload}: N={n_global}, mesh {mesh[0]}x{mesh[1]}x{mesh[2]}, "
                        f"T={T} (U={U} distinct potential rows), sigma={cfg.sigma}, "
                        f"kappa={cfg.kappa}, DefaultWithChi" + (", PME" if pme else ""),
            "inputs": f"{NBUF} trajectory frames visited ping-pong, consecutive frames differ by one "
                      "outer step of thermal motion (0.25 ps at 323 K, ~0.05 nm)",
            "parallelism": "single GPU" if world == 1 else
                           f"{world} x-slabs: slab FFT with transposes and halos stored into peer HBM "
                           "over NVLink (CUDA IPC), particles routed to their slab every step inside "
                           "the timed loop"}


NBUF = 6
CHUNKS = 8


class OracleCycle:
    """The CPU restatement of the cycle (threaded C CIC loops + scipy.fft on all cores); the mesh
    state is allocated once, like the reference's initialize_pm."""

    def __init__(self, cfg, dtype):
        import copy
        from oracle import field_oracle as fo
        from oracle.hamiltonian_oracle import OracleHamiltonian
        self.fo, self.dtype = fo, dtype
        self.cfg = copy.deepcopy(cfg)
        self.h = OracleHamiltonian(self.cfg)
        self.st = fo.FieldState(self.cfg, dtype)

    def __call__(self, pos, typ, q):
        fo, cfg = self.fo, self.cfg
        pos = np.ascontiguousarray(pos, dtype=self.dtype)
        t0 = time.perf_counter()
        fo.update_field(self.st, self.h, pos, typ, cfg, workers=-1, mt=True)
        f = fo.compute_field_force(self.st, pos, typ, cfg.n_types, mt=True)
        ef = None
        if q is not None:
            ef = fo.update_field_force_q(self.st, self.h, np.asarray(q, dtype=self.dtype), pos, cfg,
                                         workers=-1, mt=True)
        return f, ef, time.perf_counter() - t0


def oracle_cycle(cfg, pos, typ, q, dtype):
    return OracleCycle(cfg, dtype)(pos, typ, q)


def run_reference(args, rank, world):
    """CPU port of the reference path (oracle/), all host threads; rank 0 only."""
    if rank != 0:
        return
    from hymd_b200.synthetic import SPECS, make_system
    dtype = np.float32 if args.dtype == "f32" else np.float64
    spec = SPECS[args.workload]
    n_full, mesh_full = args.n or spec["n"], args.mesh or spec["mesh"]
    copies = world if args.scaling == "weak" else 1
    # The arm runs the workload the GPU arm names, at full size, for every one of the K + W cycles
    # (measured on the GPU box: 0.44 us per particle-step on 16 cores, i.e. ~4.4 s per C4 cycle and
    # ~100 s for the driver's 20 + 5 cycles).  Only a configuration whose K + W full cycles would
    # exceed 20 minutes (C5: ~50 s per cycle) is sampled: same density and mesh spacing, box
    # halved per axis; the line then says so (same_config false).
    total = args.steps + args.warmup
    n, mesh, shrink = n_full, mesh_full, 0
    while n * copies * total * 0.6e-6 > 1200.0 and mesh % 2 == 0 and mesh >= 32:
        n, mesh, shrink = n // 8, mesh // 2, shrink + 1
    if copies == 1:
        sysm = make_system(args.workload, dtype=dtype, n=n, mesh=mesh)
        pos, typ, q = sysm.positions, sysm.types, sysm.charges
    else:   # weak scaling: the global system is `world` boxes stacked along x
        parts = [make_system(args.workload, dtype=dtype, n=n, mesh=mesh, x_copies=copies, x_index=r)
                 for r in range(copies)]
        sysm = parts[0]
        pos = np.concatenate([p.positions for p in parts])
        typ = np.concatenate([p.types for p in parts])
        q = None if sysm.charges is None else np.concatenate([p.charges for p in parts])
    cfg = sysm.config
    n_run = len(pos)
    cores = os.cpu_count() or 1
    oc = OracleCycle(cfg, dtype)
    for _ in range(args.warmup):
        oc(pos, typ, q)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oc(pos, typ, q)
    dt = time.perf_counter() - t0
    value = n_run * args.steps / dt
    sample = (f"every step = one full cycle of {args.workload} (N={n_run}, mesh {mesh}), nothing sampled"
              if shrink == 0 else
              f"{args.workload} shrunk {2 ** shrink}x per axis at equal density and mesh spacing "
              f"(N={n_run}, {mesh}^3 instead of N={n_full * copies}, {mesh_full}^3)")
    config = workload_config(args, world, cfg, n_run, q is not None)
    config["same_config"] = shrink == 0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": config,
        "ns_per_day": args.steps / dt * 86400.0 * PS_PER_CYCLE / 1000.0,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample,
                         "note": "CPU restatement of field.py + pmesh primitives (C CIC loops "
                                 "on all threads + scipy.fft workers=-1); the reference's own "
                                 "pmesh/PFFT/mpi4py stack is not installable in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


PARITY_TOL = {"f32": 1e-5, "f64": 1e-10}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (hymd_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from hymd_b200 import field as F
    from hymd_b200.hamiltonian import get_hamiltonian

    np_dtype = np.float32 if args.dtype == "f32" else np.float64
    t_dtype = torch.float32 if args.dtype == "f32" else torch.float64
    b = 4 if args.dtype == "f32" else 8
    sysm, presharded = build_system(args, world, rank)
    cfg = sysm.config
    mesh = [int(x) for x in np.full(3, cfg.mesh_size)]
    T = cfg.n_types
    N = len(sysm.positions)
    if presharded and world > 1:
        nt = torch.tensor([N], dtype=torch.int64, device="cuda")
        dist.all_reduce(nt)
        N = int(nt.item())
    ham = get_hamiltonian(cfg)
    pm, fl, ecl, cl = F.initialize_pm(None, cfg)
    phi, phi_fourier, force_mesh, v_ext_fourier, v_ext, phi_transfer, phi_laplacian = fl
    phi_q, phi_q_fourier, psi, elec_field = ecl
    layouts = [pm.decompose(None) for _ in range(T)]
    pme = sysm.charges is not None

    # this rank's particles (slab along x); single GPU: all of them.  gid = index in the global
    # system (strong scaling: the generator's order; weak: box r holds [r*n, (r+1)*n))
    pos_h, typ_h, q_h = sysm.positions, sysm.types, sysm.charges
    vel_h = sysm.velocities
    gid_h = np.arange(len(pos_h), dtype=np.int64) + (rank * len(pos_h) if presharded else 0)   # equal shares
    if world > 1 and not presharded:
        L = float(cfg.box_size[0])
        cell = np.floor(pos_h[:, 0].astype(np.float64) * mesh[0] / L).astype(np.int64) % mesh[0]
        mine = (cell // (mesh[0] // world)) == rank
        pos_h, typ_h, vel_h, gid_h = pos_h[mine], typ_h[mine], vel_h[mine], gid_h[mine]
        q_h = None if q_h is None else q_h[mine]
    if not args.no_dd:
        # what main.py does once at start-up (and every config.domain_decomposition steps) when
        # the option is set: all per-particle arrays come back permuted identically
        extra = (vel_h, typ_h, gid_h) if q_h is None else (vel_h, typ_h, gid_h, q_h)
        out = F.domain_decomposition(pos_h, pm, *extra)
        pos_h, vel_h, typ_h, gid_h = out[0], out[1], out[2], out[3]
        q_h = None if q_h is None else out[4]
    n_loc = len(pos_h)
    dev = pm.device
    # MD-like input stream: step k sees the positions of step k-1 displaced by one outer step of
    # thermal motion (v * respa_inner * time_step, ~0.05 nm = 12 % of a cell), so every step
    # re-bins genuinely different coordinates.  NBUF trajectory frames, visited ping-pong.  With
    # several GPUs nothing keeps a particle inside its rank's slab: frames 1.. hold particles that
    # crossed a slab face, and every cycle routes them to the slab owner and their forces back.
    L = np.asarray(cfg.box_size, dtype=np.float64)
    frames_h = []
    for k in range(NBUF):
        f = np.mod(pos_h.astype(np.float64) + k * PS_PER_CYCLE * vel_h.astype(np.float64), L)
        f = f.astype(np_dtype)
        f[f >= L.astype(np_dtype)] = 0
        if world > 1 and CLIP_TO_SLAB:
            lo = rank * (mesh[0] // world) * L[0] / mesh[0]
            hi = (rank + 1) * (mesh[0] // world) * L[0] / mesh[0]
            f[:, 0] = np.clip(f[:, 0], np.nextafter(np_dtype(lo), np_dtype(hi)) if lo > 0 else 0,
                              np.nextafter(np_dtype(hi), np_dtype(lo)))
        frames_h.append(np.ascontiguousarray(f))
    order = list(range(NBUF)) + list(range(NBUF - 2, 0, -1))
    frames_d = [torch.as_tensor(f, dtype=t_dtype, device=dev) for f in frames_h]
    typ_d = torch.as_tensor(typ_h.astype(np.int32), device=dev)
    q_d = None if q_h is None else torch.as_tensor(q_h, dtype=t_dtype, device=dev)
    force_d = torch.zeros((n_loc, 3), dtype=t_dtype, device=dev)
    eforce_d = torch.zeros((n_loc, 3), dtype=t_dtype, device=dev) if pme else None
    counter = {"dev": 0, "host": 0}

    def cycle(pos, typ, q, force, eforce):
        F.update_field(phi, phi_laplacian, phi_transfer, layouts, force_mesh, ham, pm, pos, typ,
                       cfg, v_ext, phi_fourier, v_ext_fourier, cfg.m, compute_potential=False)
        F.compute_field_force(layouts, pos, force_mesh, force, typ, T)
        if pme:
            F.update_field_force_q(q, phi_q, phi_q_fourier, psi, None, None, elec_field, eforce,
                                   pm.decompose(None), ham, pm, pos, cfg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between barrier+synchronize, CUDA events, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    # ---- device-resident run ---------------------------------------------------------------
    def dev_cycle():
        k = order[counter["dev"] % len(order)]
        counter["dev"] += 1
        cycle(frames_d[k], typ_d, q_d, force_d, eforce_d)

    # (small systems: a few more untimed steps, so that the graph replay of the step -- csrc/graph.cu,
    # recorded on the second sight of a call shape -- is what the timed region measures when it is enabled)
    n_warm = max(args.warmup, 3) if n_loc > (1 << 21) else max(args.warmup, 8)
    for _ in range(n_warm):
        dev_cycle()
    torch.cuda.synchronize()
    graph0 = pm.graph_stats()
    replaying = graph0["replayed"] > 0
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if not replaying:
        pm.set_timing(True)
        pm.timings()
        launches0 = pm.launch_count()
        ms_total = timed(dev_cycle, args.steps)
        launches = pm.launch_count() - launches0
        phase_ms = pm.timings()
        pm.set_timing(False)
    else:
        # phase events force the ordinary launches: the timed steps run without them, the phases are
        # measured in a second pass over the same steps
        launches0 = pm.launch_count()
        ms_total = timed(dev_cycle, args.steps)
        launches = pm.launch_count() - launches0
        pm.set_timing(True)
        pm.timings()
        timed(dev_cycle, args.steps)
        phase_ms = pm.timings()
        pm.set_timing(False)
    clocks = sampler.stop() if sampler else None
    graph1 = pm.graph_stats()
    ms_per_step = ms_total / args.steps
    N_global = N
    value = N_global * args.steps / (ms_total * 1e-3)

    # ---- parity: the GPU forces of the most-drifted frame against the CPU oracle -------------
    # Same frame on every rank (NBUF - 1 outer steps of drift: with several GPUs that is the frame
    # with the most particles away from their home slab).  The oracle runs in float64 on the same
    # (dtype-valued) coordinates, the tolerance is the north star's (1e-5 fp32 / 1e-10 fp64),
    # max|a-b| / max|b| over all particles of all ranks.
    PF = NBUF - 1
    cycle(frames_d[PF], typ_d, q_d, force_d, eforce_d)
    torch.cuda.synchronize()
    gpu_f = force_d.cpu().numpy().copy()
    gpu_ef = eforce_d.cpu().numpy().copy() if pme else None
    parity = None
    if N_global <= args.parity_max_n:
        def gather_rows(a):
            """Per-rank arrays with one row per local particle -> list of the ranks' arrays on rank 0 (NCCL
            gather of padded tensors: the ranks hold different particle counts after domain_decomposition)."""
            if a is None:
                return None
            t = torch.as_tensor(np.ascontiguousarray(a), device=dev)
            nmax = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
            counts = [torch.zeros_like(nmax) for _ in range(world)]
            dist.all_gather(counts, nmax)
            m = int(max(int(c.item()) for c in counts))
            pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            pad[: t.shape[0]] = t
            out = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
            dist.gather(pad, out, dst=0)
            if rank != 0:
                return None
            return [o[: int(c.item())].cpu().numpy() for o, c in zip(out, counts)]

        if world > 1:
            cols = [gid_h, frames_h[PF] if presharded else None, typ_h if presharded else None,
                    q_h if presharded else None, gpu_f, gpu_ef]
            cols = [gather_rows(a) for a in cols]
            gathered = None if rank != 0 else \
                [tuple(None if c is None else c[r] for c in cols) for r in range(world)]
        else:
            gathered = [(gid_h, frames_h[PF], typ_h, q_h, gpu_f, gpu_ef)]
        if rank == 0:
            if presharded:      # weak scaling: the global system is the ranks' boxes side by side
                g_pos = np.concatenate([g[1] for g in gathered])
                g_typ = np.concatenate([g[2] for g in gathered])
                g_q = None if not pme else np.concatenate([g[3] for g in gathered])
                got = np.concatenate([g[4] for g in gathered])
                got_e = None if not pme else np.concatenate([g[5] for g in gathered])
            else:               # strong scaling: frame PF of the whole system in generator order
                g_pos = np.mod(sysm.positions.astype(np.float64) +
                               PF * PS_PER_CYCLE * sysm.velocities.astype(np.float64), L).astype(np_dtype)
                g_pos[g_pos >= L.astype(np_dtype)] = 0
                g_typ, g_q = sysm.types, sysm.charges
                got = np.full((N_global, 3), np.nan)
                got_e = np.full((N_global, 3), np.nan) if pme else None
                for g in gathered:
                    got[g[0]] = g[4]
                    if pme:
                        got_e[g[0]] = g[5]
                if CLIP_TO_SLAB and world > 1:
                    raise SystemExit("parity needs unclipped frames")
            want, want_e, oracle_s = oracle_cycle(cfg, g_pos, g_typ, g_q, np.float64)
            err = rel_err(got, want) if np.isfinite(got).all() else float("inf")
            parity = {"config": args.workload, "n_compared": int(len(want)), "frame": PF,
                      "rel_err": err, "tol": PARITY_TOL[args.dtype],
                      "oracle": "oracle/field_oracle.py, float64, one cycle on rank 0 "
                                f"({oracle_s:.1f} s)", "norm": "max|a-b| / max|b| over all particles"}
            if pme:
                parity["rel_err_elec"] = rel_err(got_e, want_e)
            net = np.abs(got.sum(axis=0)).max() / max(np.abs(got).sum(), 1e-300)
            parity["net_force_fraction"] = float(net)
            parity["ok"] = bool(err < parity["tol"] and parity.get("rel_err_elec", 0.0) < parity["tol"])
    else:
        # too large for a CPU oracle cycle inside the bench: the size-independent property only
        # (momentum conservation of the spectral field forces, summed over all ranks)
        s1 = torch.tensor(np.concatenate([gpu_f.astype(np.float64).sum(axis=0),
                                          [np.abs(gpu_f).astype(np.float64).sum()]]), device=dev)
        if world > 1:
            dist.all_reduce(s1)
        s1 = s1.cpu().numpy()
        net = float(np.abs(s1[:3]).max() / max(s1[3], 1e-300))
        parity = {"config": args.workload, "n_compared": 0, "rel_err": None,
                  "note": f"N={N_global} exceeds --parity-max-n: momentum conservation only",
                  "net_force_fraction": net, "ok": bool(net < (1e-5 if args.dtype == "f32" else 1e-10))}

    # ---- end to end: host (pinned) in, host out, every step --------------------------------
    e2e = None
    if not args.no_e2e:
        frames_p = [torch.from_numpy(f).pin_memory() for f in frames_h[:4]]
        order_p = [0, 1, 2, 3, 2, 1]
        typ_p = torch.from_numpy(typ_h.astype(np.int32)).pin_memory()
        q_p = None if q_h is None else torch.from_numpy(np.ascontiguousarray(q_h)).pin_memory()
        force_p = torch.zeros((n_loc, 3), dtype=t_dtype).pin_memory()
        eforce_p = torch.zeros((n_loc, 3), dtype=t_dtype).pin_memory() if pme else None

        def host_cycle():
            k = order_p[counter["host"] % len(order_p)]
            counter["host"] += 1
            cycle(frames_p[k], typ_p, q_p, force_p, eforce_p)
            return k

        for _ in range(2):
            host_cycle()
        k = max(3, min(args.steps, 10))
        ms_e2e = timed(host_cycle, k)
        # the last host step against the same frame computed device-resident
        last = order_p[(counter["host"] - 1) % len(order_p)]
        cycle(frames_d[last], typ_d, q_d, force_d, eforce_d)
        torch.cuda.synchronize()
        if not np.isfinite(force_p.numpy()).all() or \
                not torch.equal(force_p, force_d.cpu()):
            raise SystemExit("end-to-end forces differ from the device-resident run")
        pos_p = frames_p[0]
        # positions (and charges for PME) cross PCIe every step; the types are static per-particle
        # data and are uploaded once (the API re-uploads them only when a different array is passed)
        h2d = pos_p.numel() * pos_p.element_size() + \
            (q_p.numel() * q_p.element_size() if pme else 0)
        d2h = force_p.numel() * force_p.element_size() * (2 if pme else 1)
        e2e = {"value": N_global * k / (ms_e2e * 1e-3), "unit": UNIT, "steps": k,
               "ms_per_step": ms_e2e / k, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h),
               "path": "hymd_b200.field.update_field + compute_field_force with pinned host "
                       "tensors (per rank)"}
        # what the unmodified main.py hands over: pageable numpy arrays, forces written in place
        # into a numpy array (wall clock around synchronised steps)
        force_n = np.zeros((n_loc, 3), dtype=np_dtype)
        eforce_n = np.zeros((n_loc, 3), dtype=np_dtype) if pme else None
        typ_n = typ_h.astype(np.int32)
        kn = 3
        cycle(frames_h[0], typ_n, q_h, force_n, eforce_n)
        barrier()
        t0 = time.perf_counter()
        for i in range(kn):
            cycle(frames_h[1 + i % 3], typ_n, q_h, force_n, eforce_n)
        torch.cuda.synchronize()
        dt_n = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt_n, op=dist.ReduceOp.MAX)
        e2e["numpy_pageable"] = {"value": N_global * kn / float(dt_n.item()), "unit": UNIT, "steps": kn,
                                 "ms_per_step": float(dt_n.item()) / kn * 1e3,
                                 "path": "same calls with pageable numpy arrays in and out (what "
                                         "main.py passes), wall clock, max over ranks"}
        barrier()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline / phases -----------------------------------------------------------------
    peak, peak_src = peaks()
    U = pm.status()["potential_rows"]
    loc_mesh = [mesh[0] // world, mesh[1], mesh[2]]
    alg = algorithmic_bytes(n_loc, T, U, loc_mesh, b, pme)
    traffic = ncu_traffic(args.workload, args.dtype)
    phases = {}
    for name, (ms, calls) in phase_ms.items():
        per = ms / args.steps
        ent = {"ms_per_step": per, "share": per / ms_per_step}
        if name in alg:
            ent["algorithmic_bytes"] = int(alg[name])
            ent["gbs"] = alg[name] / (per * 1e-3) / 1e9 if per > 0 else None
            ent["frac_of_peak"] = ent["gbs"] / peak if per > 0 else None
        phases[name] = ent
    paths = pm.paths()
    kernel_of = {"paint": "paint_kernel", "readout": "readout_gather_kernel",
                 "kspace": "xline_kernel" if paths["xline"] else "kspace_force_kernel"}
    if paths["plane"]:          # the (y,z) transforms are hand-written too (planefft.cu)
        kernel_of.update({"fft_fwd": "plane_r2c_kernel", "fft_inv": "plane_c2r_kernel"})
    if os.environ.get("HYMD_B200_READOUT", "g")[0] == "t":
        kernel_of["readout"] = "readout_kernel"
    own = [k for k in kernel_of if k in phases]
    dom = max(own, key=lambda k: phases[k]["ms_per_step"])
    for k in own:
        phases[k]["kernel"] = kernel_of[k]
        if traffic.get(k):
            phases[k]["ncu_dram_bytes"] = traffic[k]
    nested = sum(phases[k]["ms_per_step"] for k in ("alltoall", "halo") if k in phases)
    roofline = {"kernel": kernel_of[dom],
                "bound": "hbm", "achieved": phases[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": phases[dom]["frac_of_peak"], "traffic": traffic.get(dom),
                "algorithmic_bytes": phases[dom]["algorithmic_bytes"],
                "peak_source": peak_src,
                "timing": "CUDA events around the phase on the launch stream, averaged over the "
                          "timed steps" + (f" (multi-GPU: the transform phases include {nested:.3f} ms "
                                           "of exchange + barrier per step)" if world > 1 else ""),
                "all_hand_written": {k: {"kernel": kernel_of[k], "gbs": phases[k]["gbs"],
                                         "frac": phases[k]["frac_of_peak"]} for k in own}}
    total_alg = sum(alg.values())
    config = workload_config(args, world, cfg, N_global, pme)
    config["particle_order"] = ("generator (molecule) order" if args.no_dd else
                                "as returned by the start-up domain_decomposition call "
                                "(mesh-cell order of frame 0, main.py:274-300)")
    config["l2"] = (f"inputs larger than L2 (per-step working set {total_alg / 1e6:.0f} MB per GPU "
                    "vs 126 MB L2), no flush")
    if world > 1:
        config["routing"] = "clipped to the slab (no routing)" if CLIP_TO_SLAB else \
            "in loop: unclipped trajectories, guests routed every cycle"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": config,
        "ns_per_day": args.steps / (ms_total * 1e-3) * 86400.0 * PS_PER_CYCLE / 1000.0,
        "cycle_frac_of_hbm_roofline": total_alg / (ms_per_step * 1e-3) / 1e9 / peak,
        "parity": parity,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "paths": paths,
        "graph": {"replayed_steps": graph1["replayed"] - graph0["replayed"], "graphs": graph1["alive"],
                  "phase_timing": "second pass with the ordinary launches" if replaying else "inside the timed steps"},
        "roofline": roofline, "phases": phases,
    }
    if world == 1 and not args.no_cpu_baseline:
        # the CPU port in the run's own dtype on the same frame (one full cycle, all host threads)
        reps = 1 if N_global >= 5_000_000 else (3 if N_global >= 500_000 else 10)
        oc = OracleCycle(cfg, np_dtype)
        if reps > 1:
            oc(frames_h[0], typ_h, q_h)
        dt = sum(oc(frames_h[0], typ_h, q_h)[2] for _ in range(reps))
        line["cpu_baseline"] = {"value": N_global * reps / dt, "unit": UNIT, "cores": os.cpu_count() or 1,
                                "kind": "port",
                                "sample": f"{reps} full cycle(s) of the same workload (N={N_global}), "
                                          f"{dt:.1f} s of wall time on all host threads"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        print(f"PARITY FAILED: {parity}", file=sys.stderr, flush=True)
        return 3
    return 0


CLIP_TO_SLAB = False    # debugging aid only: keeps every frame inside the slab (no guests)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Phase timings of the field-force cycle for several kernel variants in ONE process.

    python tools/variants.py [--workload C4] [--dtype f32] [--steps 10] "A=1,B=2" "A=0" ...

Every positional argument is one variant: a comma-separated list of environment settings the
library reads when a context is created / a kernel is launched (HYMD_B200_*); "-" = defaults.
The synthetic system is generated once (that is most of bench.py's wall time), each variant gets
a fresh context, 3 warm-up + `steps` timed cycles over the same MD-like frames as bench.py.
Prints one line per variant: ms per cycle and per phase (CUDA events on the launch stream)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--mesh", type=int, default=None)
    ap.add_argument("--out", default=None, help="append JSON lines here")
    ap.add_argument("variants", nargs="*", default=["-"])
    args = ap.parse_args()

    import torch
    from hymd_b200 import field as F
    from hymd_b200.hamiltonian import get_hamiltonian
    from hymd_b200.synthetic import make_system

    np_dtype = np.float32 if args.dtype == "f32" else np.float64
    t_dtype = torch.float32 if args.dtype == "f32" else torch.float64
    sysm = make_system(args.workload, dtype=np_dtype, n=args.n, mesh=args.mesh)
    cfg = sysm.config
    T = cfg.n_types
    ham = get_hamiltonian(cfg)
    pme = sysm.charges is not None

    # start-up domain_decomposition (cell order), as bench.py
    pm0 = F.initialize_pm(None, cfg)[0]
    extra = (sysm.velocities, sysm.types) if not pme else (sysm.velocities, sysm.types, sysm.charges)
    out = F.domain_decomposition(sysm.positions, pm0, *extra)
    pos_h, vel_h, typ_h = out[0], out[1], out[2]
    q_h = out[3] if pme else None
    pm0.close()
    L = np.asarray(cfg.box_size, dtype=np.float64)
    frames = []
    for k in range(4):
        f = np.mod(pos_h.astype(np.float64) + k * 0.25 * vel_h.astype(np.float64), L).astype(np_dtype)
        f[f >= L.astype(np_dtype)] = 0
        frames.append(torch.as_tensor(f, dtype=t_dtype, device="cuda"))
    order = [0, 1, 2, 3, 2, 1]
    typ_d = torch.as_tensor(typ_h.astype(np.int32), device="cuda")
    q_d = None if q_h is None else torch.as_tensor(q_h, dtype=t_dtype, device="cuda")
    n = len(pos_h)
    force = torch.zeros((n, 3), dtype=t_dtype, device="cuda")
    eforce = torch.zeros((n, 3), dtype=t_dtype, device="cuda") if pme else None
    ref_force = None

    for var in args.variants:
        sets = {} if var == "-" else dict(kv.split("=", 1) for kv in var.split(","))
        saved = {k: os.environ.get(k) for k in sets}
        os.environ.update(sets)
        try:
            pm, fl, ecl, _ = F.initialize_pm(None, cfg)
            phi, phi_fourier, force_mesh, v_ext_fourier, v_ext, phi_transfer, phi_laplacian = fl
            phi_q, phi_q_fourier, psi, elec_field = ecl
            layouts = [pm.decompose(None) for _ in range(T)]
            step = [0]

            def cycle():
                pos = frames[order[step[0] % len(order)]]
                step[0] += 1
                F.update_field(phi, phi_laplacian, phi_transfer, layouts, force_mesh, ham, pm, pos,
                               typ_d, cfg, v_ext, phi_fourier, v_ext_fourier, cfg.m)
                F.compute_field_force(layouts, pos, force_mesh, force, typ_d, T)
                if pme:
                    F.update_field_force_q(q_d, phi_q, phi_q_fourier, psi, None, None, elec_field,
                                           eforce, pm.decompose(None), ham, pm, pos, cfg)

            for _ in range(3):
                cycle()
            torch.cuda.synchronize()
            pm.set_timing(True)
            pm.timings()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                cycle()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            ph = {k: round(v[0] / args.steps, 4) for k, v in pm.timings().items()}
            # all variants must agree on the forces of the last frame (same arithmetic)
            cur = force.clone()
            if ref_force is None:
                ref_force, dev = cur, 0.0
            else:
                dev = float((cur - ref_force).abs().max() / ref_force.abs().max())
            line = {"variant": var, "ms_per_cycle": round(ms, 4), "phases": ph,
                    "max_rel_dev_vs_first": dev, "paths": pm.paths()}
            print(json.dumps(line), flush=True)
            if args.out:
                with open(args.out, "a") as fh:
                    fh.write(json.dumps(line) + "\n")
            pm.close()
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return 0


if __name__ == "__main__":
    sys.exit(main())

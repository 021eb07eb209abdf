"""Device timing of the row-f2 kernels (bonded forces, fused kick/drift, CSVR) at C4 size.

    python tools/bench_md.py [--n 10000000] [--iters 20] [--out gpurun_out/md_bench.json]

Synthetic system: half of the particles in 20-bead random-walk chains (bond 0.47 nm; bonds between
consecutive beads, angles over consecutive triples), half solvent, particles in molecule order, box at
the HyMD density.  CUDA events on the launch stream, warm-up first.  Algorithmic bytes (fp32, b = 4):
  bonded kind k : N*(3b + 3b + 4) + terms*(16 + 8*params + 4*k)   positions in, forces out, CSR start,
                  term indices / parameters / per-particle references once
  kick+drift    : N*3b*(2 + 2 + n_forces)                          v, x read+write, forces read
  csvr          : N*3b (moments) + N*3b*2 (rescale) + N*4*2 (group ids)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hymd_b200 import thermostat as T  # noqa: E402
from hymd_b200.force import BondedTopology  # noqa: E402
from hymd_b200.md import kick_drift  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    n, L = args.n, float((args.n / 8.37) ** (1.0 / 3.0))
    rng = np.random.default_rng(1004)
    nch = n // 40
    steps = rng.normal(size=(nch, 20, 3)).astype(np.float32)
    steps *= 0.47 / np.linalg.norm(steps, axis=2, keepdims=True)
    steps[:, 0] = rng.uniform(0, L, size=(nch, 3))
    pos = np.empty((n, 3), dtype=np.float32)
    pos[:nch * 20] = np.mod(np.cumsum(steps, axis=1).reshape(-1, 3), L)
    pos[nch * 20:] = rng.uniform(0, L, size=(n - nch * 20, 3))
    first = (np.arange(nch, dtype=np.int32) * 20)[:, None]
    a2 = (first + np.arange(19, dtype=np.int32)[None, :]).ravel()
    a3 = (first + np.arange(18, dtype=np.int32)[None, :]).ravel()
    t0 = time.time()
    topo = BondedTopology(n, bonds=(a2, a2 + 1, np.full(len(a2), 0.47), np.full(len(a2), 1250.0)),
                          angles=(a3, a3 + 1, a3 + 2, np.full(len(a3), np.pi), np.full(len(a3), 25.0)))
    t_create = time.time() - t0
    box = np.array([L, L, L])
    x = torch.as_tensor(pos, device="cuda")
    v = torch.randn((n, 3), device="cuda") * 0.19
    fb, fa, ff = torch.empty_like(x), torch.empty_like(x), torch.randn_like(x)
    grp = torch.as_tensor((np.arange(n) >= nch * 20).astype(np.int32), device="cuda")
    res = {"n": n, "bonds": int(len(a2)), "angles": int(len(a3)), "topology_create_s": t_create}
    peak = None
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                               "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh).get("hbm_gbs"))
    except Exception:
        pass

    def line(name, ms, nbytes):
        res[name] = {"ms": ms, "algorithmic_MB": nbytes / 1e6, "GBps": nbytes / ms / 1e6,
                     "frac_of_measured_peak": (nbytes / ms / 1e6 / peak) if peak else None}
    ms = timed(lambda: topo.forces(2, x, box, fb), args.iters)
    line("bonds", ms, n * 28 + len(a2) * (16 + 16 + 8))
    ms = timed(lambda: topo.forces(3, x, box, fa), args.iters)
    line("angles", ms, n * 28 + len(a3) * (16 + 16 + 12))
    x2, v2 = x.clone(), v.clone()
    ms = timed(lambda: kick_drift(v2, x2, [fb, fa], 72.0, 0.01, 0.01, box), args.iters)
    line("kick_drift_2forces", ms, n * 12 * 6)
    ms = timed(lambda: kick_drift(v2, None, [fb, fa], 72.0, 0.01), args.iters)
    line("kick_2forces", ms, n * 12 * 4)
    ms = timed(lambda: kick_drift(v2, None, [ff], 72.0, 0.25, sequential=True), args.iters)
    line("kick_1force", ms, n * 12 * 3)
    v3 = v.clone()
    mom_ms = timed(lambda: T.velocity_moments(v3, grp, 0, allreduce=False), args.iters)
    line("velocity_moments", mom_ms, n * 16)

    class Cfg:
        gas_constant, mass, target_temperature, time_step, respa_inner, tau = 0.0083144621, 72.0, 323.0, 0.01, 25, 0.1
        thermostat_coupling_groups = [["P"], ["W"]]
        name_to_type_map = {"P": 0, "W": 1}
        unique_names = ["P", "W"]
        thermostat_work = 0.0
    cfg = Cfg()
    prng = np.random.default_rng(3)
    ms = timed(lambda: T.csvr_thermostat(v3, grp, cfg, prng), max(args.iters // 4, 3))
    line("csvr_2groups", ms, 2 * n * (12 + 24 + 8))
    x3, x4, v4 = x.clone(), torch.empty_like(x), v.clone()
    fused_bytes = n * 12 * 4 + len(a2) * 40 + len(a3) * 44 + n * 8
    ms = timed(lambda: topo.inner_step(x3, x4, v4, box, 72.0, 0.01, 2, 0.0, want_energies=False), args.iters)
    line("fused_inner_step", ms, fused_bytes)
    for mode, tag in ((1, "cta"), (2, "cta2"), (3, "cta3")):    # CTA-cooperative term evaluation (hymd_bonded_set_cta)
        try:
            topo.set_cta(mode)
            ms = timed(lambda: topo.forces(2, x, box, fb), args.iters)
            line("bonds_" + tag, ms, n * 28 + len(a2) * (16 + 16 + 8))
            ms = timed(lambda: topo.forces(3, x, box, fa), args.iters)
            line("angles_" + tag, ms, n * 28 + len(a3) * (16 + 16 + 12))
            ms = timed(lambda: topo.inner_step(x3, x4, v4, box, 72.0, 0.01, 2, 0.0, want_energies=False), args.iters)
            line("fused_inner_step_" + tag, ms, fused_bytes)
        except Exception as exc:   # keep the lines measured so far
            res[tag + "_error"] = repr(exc)
    topo.set_cta(0)
    try:    # single-precision bond / angle arithmetic (hymd_bonded_set_math)
        topo.set_math(True)
        ms = timed(lambda: topo.forces(2, x, box, fb), args.iters)
        line("bonds_f32math", ms, n * 28 + len(a2) * (16 + 16 + 8))
        ms = timed(lambda: topo.forces(3, x, box, fa), args.iters)
        line("angles_f32math", ms, n * 28 + len(a3) * (16 + 16 + 12))
        ms = timed(lambda: topo.inner_step(x3, x4, v4, box, 72.0, 0.01, 2, 0.0, want_energies=False), args.iters)
        line("fused_inner_step_f32math", ms, fused_bytes)
    except Exception as exc:
        res["f32math_error"] = repr(exc)
    topo.set_math(False)
    inner = res["kick_drift_2forces"]["ms"] + res["bonds"]["ms"] + res["angles"]["ms"] + res["kick_2forces"]["ms"]
    res["inner_rrespa_step_ms"] = inner
    res["launches_topology"] = topo.launch_count()
    print(json.dumps(res))
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()

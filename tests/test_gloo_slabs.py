"""Multi-process CPU tests (gloo, world_size 2 and 4) of the N > 1 path's host-side logic:

* the slab decomposition itself: ownership by x-cell, ghost-plane halo
  reduce after the paint, k layout after the forward transpose, halo fetch before the readout --
  restated on CPUs in oracle/slab_oracle.py and compared with the single-rank oracle;
* bench.py's reference arm under torchrun: rank 0 alone measures and prints ONE JSON line, the
  other ranks exit 0 without work;
* the weak-scaling generator: every rank builds its own slab of the stacked boxes.

The CUDA implementation of the same data flow is tested on GPUs in tests/test_mgpu.py."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "gloo_worker.py")


def _torchrun(nproc, script, extra, port, timeout=300):
    env = dict(os.environ)
    env.pop("CUDA_VISIBLE_DEVICES", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), script] + extra
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("nproc,mesh,pme", [(2, [16, 12, 10], True), (2, [18, 12, 10], False),
                                            (4, [16, 12, 9], True)])
def test_slab_decomposition_matches_single_rank_oracle(nproc, mesh, pme):
    r = _torchrun(nproc, WORKER, ["--mesh"] + [str(m) for m in mesh] + (["--pme"] if pme else []),
                  29581 + nproc)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "OK" in r.stdout


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    r = _torchrun(2, os.path.join(ROOT, "bench.py"),
                  ["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                   "--workload", "C1"], 29591)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_weak_scaling_generator_builds_disjoint_slabs():
    from hymd_b200.synthetic import make_system
    parts = [make_system("C1", dtype=np.float32, x_copies=2, x_index=r) for r in range(2)]
    cfg = parts[0].config
    L = float(cfg.box_size[1])
    assert list(np.full(3, cfg.mesh_size)) == [48, 24, 24] or list(cfg.mesh_size) == [48, 24, 24]
    assert np.isclose(float(cfg.box_size[0]), 2 * L)
    assert cfg.n_particles == 20000
    for r, s in enumerate(parts):
        x = s.positions[:, 0]
        assert (x >= r * L).all() and (x < (r + 1) * L + 1e-4).all()
    assert not np.array_equal(parts[0].positions[:, 1], parts[1].positions[:, 1])

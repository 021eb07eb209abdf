"""Slab-sharded cycle: the same parity gates as tests/test_gpu_parity.py, through the slab
pipeline (2-D cuFFT per plane + pack/all-to-all + x transform, halo reduce/fetch).

* one GPU, HYMD_B200_FORCE_SLAB=1: the slab pipeline without the exchange
* two GPUs (skipped when the box has one): torchrun, NCCL over NVLink"""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
WORKER = os.path.join(ROOT, "tests", "mgpu_worker.py")


def _run(nproc, extra, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    if nproc == 1:
        cmd = [sys.executable, WORKER] + extra
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1", "--master-port", "29571",
               WORKER] + extra
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "OK" in r.stdout


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("mesh", [[32, 24, 40], [9, 12, 10]])
def test_slab_pipeline_single_gpu(dtype, mesh):
    _run(1, ["--dtype", dtype, "--pme", "--mesh"] + [str(m) for m in mesh] + ["--particles", "6000"],
         {"HYMD_B200_FORCE_SLAB": "1"})


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_slab_pipeline_plane_kernels_single_gpu(dtype):
    """Power-of-two square (y,z) planes: the one-pass plane transforms inside the unfused slab
    pipeline (cuFFT only along x)."""
    _run(1, ["--dtype", dtype, "--pme", "--mesh", "24", "32", "32", "--particles", "6000"],
         {"HYMD_B200_FORCE_SLAB": "1"})


def _need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_two_slabs_match_oracle(dtype):
    _need(2)
    _run(2, ["--dtype", dtype, "--pme"])


@pytest.mark.parametrize("mesh", [[32, 32, 32], [16, 64, 64]])
def test_two_slabs_plane_kernels(mesh):
    """Plane transforms + fused x-line kernel on two slabs."""
    _need(2)
    _run(2, ["--dtype", "f32", "--pme", "--mesh"] + [str(m) for m in mesh] + ["--particles", "20000"])


def test_two_slabs_unfused_path():
    _need(2)
    _run(2, ["--dtype", "f64", "--pme"], {"HYMD_B200_NO_FUSED": "1"})


def test_two_slabs_odd_planes():
    _need(2)
    _run(2, ["--dtype", "f64", "--mesh", "18", "12", "10", "--particles", "3000"])


def test_four_slabs_match_oracle():
    _need(4)
    _run(4, ["--dtype", "f64", "--pme", "--mesh", "32", "24", "40"])


def test_migration_rehomes_particles():
    _need(2)
    _run(2, ["--dtype", "f64", "--pme", "--migrate"])


@pytest.mark.parametrize("nproc", [2, 4])
@pytest.mark.parametrize("pme", [False, True])
def test_guests_are_routed_to_their_slab_and_back(pme, nproc):
    """Particles that are not on the rank owning their slab and NO domain_decomposition (what a molecule
    straddling a slab face looks like): every field call routes them to the owners' guest inboxes over
    NVLink and their forces back, like pmesh's Layout.exchange; results equal the oracle.  (The same data
    flow runs on one GPU in tests/test_gpu_virtual_slabs.py.)"""
    _need(nproc)
    _run(nproc, ["--dtype", "f64", "--route", "--mesh", "32", "32", "32"] + (["--pme"] if pme else []))


def test_nccl_barrier_fallback():
    """HYMD_B200_NCCL_BARRIER=1: the round-1 all-reduce barrier instead of flags in peer memory."""
    _need(2)
    _run(2, ["--dtype", "f32", "--pme", "--route", "--mesh", "32", "32", "32"], {"HYMD_B200_NCCL_BARRIER": "1"})


@pytest.mark.parametrize("mode", ["fused", "kernels"])
def test_exchange_modes_over_nvlink(mode):
    """HYMD_B200_EXCHANGE: the non-default transposes (remote stores from the kernels' epilogues, pack / unpack
    push kernels) over real peer memory."""
    _need(2)
    _run(2, ["--dtype", "f32", "--pme", "--mesh", "32", "32", "32"], {"HYMD_B200_EXCHANGE": mode})

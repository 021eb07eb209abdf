"""Dry run of the row-f2 host layer and of the GPU test bodies on the CPU.

tests/native/host_check.cpp exports the row-f2 entry points of include/hymd_b200.h with the same
names and signatures, implemented as plain loops over the same ``__host__ __device__`` per-particle
functions the sm_100a kernels call.  Here ``hymd_b200._lib.load`` is pointed at that shim and the few
``torch.cuda`` touch points are neutralised, so hymd_b200/force.py, thermostat.py and md.py -- and the
bodies of tests/test_zgpu_md.py -- run on CPU tensors in the GPU-less container.  What this does NOT
cover is exactly what the ``-m gpu`` run adds: launch geometry, warp/block reductions, device memory.
It is test infrastructure only: the product never loads the shim."""
import contextlib
import ctypes
import os
import shutil
import subprocess
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
F2 = ["hymd_bonded_set_math", "hymd_bonded_set_cta", "hymd_bonded_inner_step", "hymd_bonded_create", "hymd_bonded_destroy", "hymd_bonded_forces", "hymd_bonded_launch_count",
      "hymd_md_kick_drift", "hymd_velocity_moments", "hymd_velocity_moments_scratch_doubles",
      "hymd_csvr_apply", "hymd_cancel_com",
      "hymd_bonded_set_last", "hymd_bonded_dipoles", "hymd_dipole_redistribute"]


@pytest.fixture()
def emulated(tmp_path_factory, monkeypatch):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("native") / "libhost_check.so")
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "native", "host_check.cpp")], check=True)
    host = ctypes.CDLL(so)
    from hymd_b200 import _lib, force
    real = _lib.load()

    class Shim:
        def __getattr__(self, name):
            return getattr(real, name)
    shim = Shim()
    for name in F2:
        fn = getattr(host, name)
        fn.argtypes = getattr(real, name).argtypes
        fn.restype = getattr(real, name).restype
        setattr(shim, name, fn)
    monkeypatch.setattr(_lib, "load", lambda: shim)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(force, "_default_device", lambda: "cpu")
    force._cache.clear()
    import test_zgpu_md as g
    monkeypatch.setattr(g, "DEVICE", "cpu")
    yield g
    force._cache.clear()


@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_bonded(emulated, real):
    emulated.test_bonded_forces_match_oracle(real)


@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_cbt_dihedrals(emulated, real):
    emulated.test_cbt_dihedrals_dipoles_and_redistribution_match_oracle(real)


@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_respa_md_with_cbt(emulated, real):
    emulated.test_respa_md_with_cbt_dihedrals_fused_equals_separate_launches(real)


def test_bonded_kats_and_edges(emulated):
    emulated.test_reference_kats_on_the_device_and_numpy_interface()
    emulated.test_bonded_edge_cases()


@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_kick_drift(emulated, real):
    emulated.test_kick_drift_matches_numpy(real)


@pytest.mark.parametrize("remove", [False, True])
@pytest.mark.parametrize("case,groups", [("all", []), ("abcd", [["A"], ["B"], ["C"], ["D"]]),
                                         ("abc_d", [["A", "B", "C"], ["D"]])])
def test_csvr(emulated, case, groups, remove):
    emulated.test_csvr_matches_reference_golden(case, groups, remove)


def test_csvr_interfaces_and_md(emulated):
    emulated.test_csvr_larger_system_numpy_interface_and_cancel_com()
    emulated.test_respa_md_with_bonds_conserves_energy_like_the_oracle(True)
    emulated.test_respa_md_with_bonds_conserves_energy_like_the_oracle(False)


@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_fused_inner_step(emulated, real):
    emulated.test_fused_inner_step_equals_separate_launches(real)


@pytest.mark.parametrize("tile", [128, 256, 512])
@pytest.mark.parametrize("real", [np.float32, np.float64])
def test_cta_cooperative(emulated, real, tile, monkeypatch):
    emulated.test_cta_cooperative_evaluation_matches_the_per_particle_result(real, tile, monkeypatch)


def test_f32_math_fused_step(emulated):
    emulated.test_f32_math_inner_step_stays_within_the_fp32_tolerance()


def test_fused_step_edge_cases(emulated):
    emulated.test_fused_step_edge_cases()

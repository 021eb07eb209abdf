import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle_clib():
    """Build oracle/_build/libcic_oracle.so if missing (gcc only; test infrastructure)."""
    import subprocess
    so = os.path.join(ROOT, "oracle", "_build", "libcic_oracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    yield


def make_config(names, n_particles, mesh, box, sigma=0.5, kappa=0.05, hamiltonian="DefaultWithChi",
                chi=(), dtype=np.float64, coulombtype=None, dielectric_const=None, m=None):
    from hymd_b200.config import Chi, Config
    cfg = Config(mesh_size=mesh, sigma=sigma, kappa=kappa, box_size=box, hamiltonian=hamiltonian,
                 chi=[Chi(*c) for c in chi], dtype=np.dtype(dtype), coulombtype=coulombtype,
                 dielectric_const=dielectric_const, m=list(m) if m is not None else [])
    cfg.finalize(names, n_particles=n_particles)
    return cfg

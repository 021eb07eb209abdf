"""One rank of the CPU (gloo) check of the multi-rank thermostat logic (row f2): every rank holds a
slice of the particles, runs hymd_b200.thermostat through the host shim of tests/native/host_check.cpp
(the C ABI's names on host memory) and the result is compared with the single-rank oracle on all
particles.  What this covers is the host layer's use of torch.distributed: the all-reduce of the group
sizes (degrees of freedom of the chi-squared draw) and of the velocity moments.
Launched by tests/test_gloo_md.py through torch.distributed.run."""
import contextlib
import ctypes
import os
import subprocess
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class Mock:
    def __init__(self, x):
        self.x, self.i, self.args = list(x), 0, []

    def __call__(self, *args):
        self.args.append(args)
        self.i += 1
        return self.x[self.i - 1]


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
    rank, P = dist.get_rank(), dist.get_world_size()
    so = sys.argv[1]
    host = ctypes.CDLL(so)
    from hymd_b200 import _lib
    from hymd_b200 import thermostat as T
    real = _lib.load()

    class Shim:
        def __getattr__(self, name):
            return getattr(real, name)
    shim = Shim()
    for name in ("hymd_velocity_moments", "hymd_velocity_moments_scratch_doubles", "hymd_csvr_apply",
                 "hymd_cancel_com"):
        fn = getattr(host, name)
        fn.argtypes, fn.restype = getattr(real, name).argtypes, getattr(real, name).restype
        setattr(shim, name, fn)
    _lib.load = lambda: shim
    torch.cuda.is_available = lambda: True
    torch.cuda.current_stream = lambda device=None: types.SimpleNamespace(cuda_stream=0)
    torch.Tensor.is_cuda = property(lambda self: True)

    from oracle import thermostat_oracle as to
    TG = np.load(os.path.join(ROOT, "tests", "golden", "thermostat_golden.npz"))
    names, v0 = TG["rand/names"], TG["rand/v0"]
    mass, gas, T0, dt, inner, tau = TG["rand/params"]
    n = len(names)
    cut = [0, n // 3, n] if P == 2 else np.linspace(0, n, P + 1).astype(int)       # uneven split
    mine = slice(cut[rank], cut[rank + 1])

    class Cfg:
        pass
    cfg = Cfg()
    cfg.mass, cfg.gas_constant, cfg.target_temperature = float(mass), float(gas), float(T0)
    cfg.time_step, cfg.respa_inner, cfg.tau = float(dt), int(inner), float(tau)
    cfg.unique_names = ["A", "B", "W"]
    cfg.name_to_type_map = {"A": 0, "B": 1, "W": 2}
    cfg.thermostat_coupling_groups = [["A", "B"], ["W"]]
    cfg.thermostat_work = 0.0
    cfg.n_particles = n
    v = torch.tensor(v0[mine])
    chi2 = Mock(TG["rand/chi2"])
    T.csvr_thermostat(v, names[mine], cfg, None, random_gaussian=Mock(TG["rand/gauss"]), random_chi_squared=chi2)
    # degrees of freedom come from the GLOBAL group sizes
    n_ab, n_w = int(np.sum(names != b"W")), int(np.sum(names == b"W"))
    assert [a[1] for a in chi2.args] == [3 * n_ab - 1, 3 * n_w - 1], chi2.args
    ref = v0.copy()
    grp = np.where(names == b"W", 1, 0).astype(np.int32)
    work = to.csvr_thermostat(ref, grp, 2, mass=mass, gas_constant=gas, target_temperature=T0, time_step=dt,
                              respa_inner=int(inner), tau=tau, draws=list(zip(TG["rand/gauss"], TG["rand/chi2"])))
    np.testing.assert_allclose(v.numpy(), ref[mine], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(v.numpy(), TG["rand/v"][mine], rtol=1e-11, atol=1e-13)
    assert abs(float(cfg.thermostat_work) - work) <= 1e-9 * abs(work)
    # cancel_com_momentum and the kinetic energy use global sums as well
    v = torch.tensor(v0[mine])
    T.cancel_com_momentum(v, cfg)
    np.testing.assert_allclose(v.numpy(), to.cancel_com_momentum(v0.copy(), n)[mine], rtol=1e-12, atol=1e-14)
    ke = float(T.kinetic_energy(torch.tensor(v0[mine]), mass))
    assert abs(ke - 0.5 * mass * np.sum(v0 ** 2)) <= 1e-10 * ke
    dist.barrier()
    if rank == 0:
        print("OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

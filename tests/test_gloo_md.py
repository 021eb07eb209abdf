"""World-size-2 and -3 CPU (gloo) tests of the row-f2 host layer: hymd_b200.thermostat with the particles
split unevenly over the ranks must reproduce the single-rank oracle (and the reference's own output in
tests/golden/thermostat_golden.npz) on every rank's slice.  See tests/gloo_md_worker.py."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT
from test_gloo_slabs import _torchrun

WORKER = os.path.join(ROOT, "tests", "gloo_md_worker.py")


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("native") / "libhost_check.so")
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "native", "host_check.cpp")], check=True)
    return so


@pytest.mark.parametrize("nproc", [2, 3])
def test_thermostat_over_ranks_matches_single_rank(host_lib, nproc):
    r = _torchrun(nproc, WORKER, [host_lib], 29611 + nproc)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "OK" in r.stdout

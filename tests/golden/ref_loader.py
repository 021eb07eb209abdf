"""Import the REFERENCE's own Python modules in the build container (golden-vector generation only).

``import hymd`` fails here: ``mpi4py``, ``h5py``, ``pmesh``, ``tomli`` and the f2py module
``force_kernels`` are not installed.  None of them is needed to *execute*
the pure-Python arithmetic of ``hymd/force.py`` (the ``*__plain`` kernels and ``prepare_bonds``),
``hymd/thermostat.py``, ``hymd/field.py``, ``hymd/pressure.py`` and ``hymd/input_parser.py`` on a
single rank, so this loader

* registers inert stand-ins for the missing third-party modules (a one-rank ``MPI.COMM_WORLD`` whose
  ``allreduce`` is the identity; ``tomli`` -> stdlib ``tomllib``; empty ``h5py`` / ``force_kernels``),
* registers ``oracle/pmesh_standin.py`` (the numpy restatement of the pmesh primitives) as ``pmesh.pm``,
* creates the package ``hymd`` with ``__path__ = [/root/reference/hymd]`` WITHOUT running its
  ``__init__`` (which imports ``main``), so ``import hymd.field`` etc. execute the reference's
  unmodified source files from where they lie.

Nothing here is imported by the product, the tests or the bench: only the ``make_*_golden.py`` scripts
use it, and only in the container that has ``/root/reference``.  The vectors they write are committed.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))


class _OneRankComm:
    """mpi4py.MPI.Intracomm for a single rank."""

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    rank = 0
    size = 1

    def allreduce(self, x, op=None):
        return x

    def allgather(self, x):
        return [x]

    def gather(self, x, root=0):
        return [x]

    def bcast(self, x, root=0):
        return x

    def reduce(self, x, op=None, root=0):
        return x

    def Barrier(self):
        pass

    barrier = Barrier

    def Allreduce(self, send, recv, op=None):
        if isinstance(send, (list, tuple)):
            send = send[0]
        out = recv[0] if isinstance(recv, (list, tuple)) else recv
        np.copyto(out, send)


def install_stubs():
    if "hymd" in sys.modules and getattr(sys.modules["hymd"], "_ref_loader", False):
        return
    if not os.path.isdir(os.path.join(REF_ROOT, "hymd")):
        raise RuntimeError("/root/reference is not available: golden vectors can only be regenerated "
                           "in the build container")
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    # numpy 2 removed aliases the reference (written for numpy 1.x) still uses
    if not hasattr(np, "string_"):
        np.string_ = np.bytes_
    if not hasattr(np, "VisibleDeprecationWarning"):
        np.VisibleDeprecationWarning = DeprecationWarning

    mpi4py = types.ModuleType("mpi4py")
    MPI = types.ModuleType("mpi4py.MPI")
    MPI.Intracomm = _OneRankComm
    MPI.Comm = _OneRankComm
    MPI.COMM_WORLD = _OneRankComm()
    MPI.SUM, MPI.MAX, MPI.MIN = "SUM", "MAX", "MIN"
    MPI.DOUBLE = "DOUBLE"
    MPI.Wtime = lambda: 0.0
    mpi4py.MPI = MPI
    sys.modules["mpi4py"] = mpi4py
    sys.modules["mpi4py.MPI"] = MPI

    import tomllib
    sys.modules.setdefault("tomli", tomllib)
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    fk = types.ModuleType("force_kernels")
    for name in ("cbf", "caf", "cdf", "cbf_d", "caf_d", "cdf_d"):
        setattr(fk, name, None)
    sys.modules["force_kernels"] = fk

    from oracle import pmesh_standin
    pmesh = types.ModuleType("pmesh")
    pmesh.pm = pmesh_standin
    pmesh.ParticleMesh = pmesh_standin.ParticleMesh
    sys.modules["pmesh"] = pmesh
    sys.modules["pmesh.pm"] = pmesh_standin

    pkg = types.ModuleType("hymd")
    pkg.__path__ = [os.path.join(REF_ROOT, "hymd")]
    pkg._ref_loader = True
    sys.modules["hymd"] = pkg


def ref(module: str):
    """``ref("force")`` -> the reference's ``hymd/force.py`` executed from /root/reference."""
    install_stubs()
    return importlib.import_module("hymd." + module)


def conftest_fixture(name: str):
    """Run one fixture function of the reference's ``test/conftest.py`` (its module-level imports of
    mpi4py / h5py are satisfied by the stubs; the pytest decorator is stripped by calling the wrapped
    function)."""
    install_stubs()
    import ast
    path = os.path.join(REF_ROOT, "test", "conftest.py")
    with open(path) as fh:
        tree = ast.parse(fh.read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np, "collections": __import__("collections")}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]()
    raise KeyError(name)

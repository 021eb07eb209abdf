"""Device-resident stand-ins for the pmesh objects HyMD threads through ``hymd/field.py``.

``main.py`` only ever (a) creates these through ``initialize_pm``, (b) passes them back into
the ``field.py`` functions and (c) calls a handful of methods on them directly:
``pm.decompose(pos[, smoothing])`` -> layout with ``get_exchange_cost()`` / ``exchange()``
(``main.py:332-334, 977-980, 1304``; ``field.py:1165-1178``), ``pm.np`` (``main.py:252``),
``pm.create(...)`` (``main.py:488-498``) and ``field.csum()`` (``field.py:693``).  Mesh fields
are therefore thin handles onto buffers owned by the CUDA context (``include/hymd_b200.h``);
``.value`` exposes them as torch tensors without copying.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, _world
from .hamiltonian import affine_parameters, get_hamiltonian


class _CudaView:
    """Minimal ``__cuda_array_interface__`` carrier for context-owned memory."""

    def __init__(self, ptr, shape, strides_bytes, typestr, owner):
        self._owner = owner  # keeps the context alive while a view exists
        self.__cuda_array_interface__ = {
            "shape": tuple(int(s) for s in shape),
            "strides": tuple(int(s) for s in strides_bytes),
            "typestr": typestr,
            "data": (int(ptr), False),
            "version": 2,
        }


class Layout:
    """What ``pm.decompose`` returns.  Binning happens inside the field calls (one counting
    sort for all types), so the layout itself carries no routing table."""

    def __init__(self, pm, n, smoothing=None):
        self.pm = pm
        self.n = int(n)
        self.smoothing = smoothing

    def get_exchange_cost(self):
        """pmesh's ``Layout.get_exchange_cost`` as ``main.py:1304-1312`` reads it: entry ``r`` = number of
        particles rank ``r`` ships to other ranks per field call, here the guests its last binning pass routed
        to other slabs (all types together: one sort serves every type's layout).  Collective."""
        return self.pm.exchange_cost()

    def exchange(self, *arrays):
        if self.pm.world_size == 1:
            return arrays if len(arrays) != 1 else arrays[0]
        return self.pm.migrate(*arrays)


class MeshField:
    """Handle onto one context-owned mesh buffer (real or complex)."""

    def __init__(self, pm, field_id, t=0, d=0, kind="real"):
        self.pm, self.field_id, self.t, self.d, self.kind = pm, field_id, t, d, kind

    def _materialize(self):
        fid = self.field_id
        if fid == _lib.FIELD_PHI_LAPLACIAN:
            self.pm.laplacian()
            return
        self.pm._materialize(
            want_phi=fid == _lib.FIELD_PHI, want_phi_fourier=fid == _lib.FIELD_PHI_FOURIER,
            want_v_ext=fid == _lib.FIELD_V_EXT,
            want_psi=fid in (_lib.FIELD_PSI, _lib.FIELD_PHI_Q_FOURIER))

    @property
    def value(self) -> torch.Tensor:
        """Zero-copy torch view of the logical (unpadded) extent of the field."""
        self._materialize()
        return self.pm._view(self.field_id, self.t, self.d, self.kind)

    @property
    def shape(self):
        return tuple(self.value.shape)

    @property
    def dtype(self):
        return self.value.dtype

    def csum(self):
        """Global sum (``RealField.csum``, ``field.py:693``)."""
        s = self.value.sum(dtype=torch.complex128 if self.kind == "complex" else torch.float64)
        return self.pm._allreduce_scalar(s)

    def cnorm(self):
        v = self.value
        s = (v.real.double() ** 2 + v.imag.double() ** 2).sum() if self.kind == "complex" \
            else (v.double() ** 2).sum()
        return self.pm._allreduce_scalar(s)

    def __array__(self, dtype=None, copy=None):
        a = self.value.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return f"MeshField(id={self.field_id}, t={self.t}, d={self.d}, kind={self.kind})"


class UnusedField:
    """Placeholder for pmesh buffers the reference allocates as scratch for its own algorithm
    (``v_ext_fourier[4]``, ``phi_transfer[3]``: ``field.py:55-57``); the fused kernels need no
    such scratch, so touching one is an error rather than a silent zero."""

    def __init__(self, name):
        self.name = name

    @property
    def value(self):
        raise RuntimeError(f"{self.name} is reference-internal scratch and is not materialized "
                           "by hymd_b200 (the fused k-space kernel does not use it)")


class ParticleMesh:
    """Stand-in for ``pmesh.pm.ParticleMesh(Nmesh, BoxSize, dtype, comm)`` (``field.py:45-47``)
    that owns one CUDA context (one slab of the mesh) on the current device."""

    def __init__(self, Nmesh, BoxSize, dtype="f4", comm=None, config=None, hamiltonian=None):
        if not torch.cuda.is_available():
            raise _lib.HymdError("hymd_b200 needs a CUDA device (there is no CPU fallback)")
        if config is None:
            raise ValueError("hymd_b200.pm.ParticleMesh needs config= (n_types, sigma, chi, ...)")
        self.lib = _lib.load()
        self.Nmesh = np.full(3, Nmesh).astype(np.int64)
        self.BoxSize = np.asarray(BoxSize, dtype=np.float64).reshape(3).copy()
        dt = np.dtype(dtype)
        if dt not in (np.dtype("f4"), np.dtype("f8")):
            raise ValueError(f"unsupported mesh dtype {dtype!r}")
        self.np_dtype = dt
        self.dtype = torch.float64 if dt == np.dtype("f8") else torch.float32
        self.cdtype = torch.complex128 if dt == np.dtype("f8") else torch.complex64
        self.comm = comm
        self.world = _world.current()     # torch.distributed ranks, virtual (in-process) ranks or one GPU
        self.world_size, self.rank = self.world.size, self.world.rank
        self.np = (self.world_size, 1)          # processor mesh, logged at main.py:252
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_types = int(config.n_types)
        self.gpe = getattr(config, "coulombtype", None) == "PIC_Spectral_GPE"
        self.pme = getattr(config, "coulombtype", None) == "PIC_Spectral" or self.gpe
        self._config = config
        self._ctx = ctypes.c_void_p()
        self._interaction_key = None
        self._sort_key = None
        self._sort_types_key = None
        self._order_types_key = None
        self._keep = (None, None, None)
        self._n_local = 0
        self._sorted_with_charges = False
        if hamiltonian is None:
            hamiltonian = get_hamiltonian(config) if _has_density_params(config) else None
        cfg = self._make_config(config, hamiltonian)
        nccl_id = self.world.comm_id(self.lib, self.device) if self.world_size > 1 else None
        _lib.check(self.lib.hymd_ctx_create(ctypes.byref(cfg), nccl_id, ctypes.byref(self._ctx)))

    # ---- context configuration ---------------------------------------------------------
    def _interaction(self, config, hamiltonian):
        T = self.n_types
        if hamiltonian is not None:
            A, c = affine_parameters(hamiltonian, T)
        else:  # parameters not known yet (set by the first update_field)
            A, c = np.zeros((T, T)), np.zeros(T)
        m = list(getattr(config, "m", None) or [1.0] * T)
        conv = 0.0
        if self.pme and getattr(config, "dielectric_const", None):
            conv = float(config.coulomb_constant) / float(config.dielectric_const)
        return A, c, np.asarray(m, dtype=np.float64), float(config.sigma), conv

    def _make_config(self, config, hamiltonian):
        A, c, m, sigma, conv = self._interaction(config, hamiltonian)
        cfg = _lib.HymdConfig()
        cfg.struct_size = ctypes.sizeof(_lib.HymdConfig)
        cfg.dtype = _lib.F64 if self.np_dtype == np.dtype("f8") else _lib.F32
        for a in range(3):
            cfg.mesh[a] = int(self.Nmesh[a])
            cfg.box[a] = float(self.BoxSize[a])
        cfg.n_types = self.n_types
        cfg.world_size, cfg.rank = self.world_size, self.rank
        cfg.pme = 1 if self.pme else 0
        cfg.sigma, cfg.elec_conversion = sigma, conv
        T = self.n_types
        for i in range(T):
            cfg.c[i] = c[i]
            cfg.m[i] = m[i]
            for j in range(T):
                cfg.A[i * T + j] = A[i, j]
        self._interaction_key = (A.tobytes(), c.tobytes(), m.tobytes(), sigma, conv)
        return cfg

    def sync_interaction(self, hamiltonian, config, m=None):
        """Push (A, c, m, sigma, k_e/eps) to the context if they changed since the last call
        (barostat runs change rho0 / a; ``update_field`` receives ``m`` explicitly)."""
        # Probing the functional (T^2 Python calls, affine_parameters) every MD step would cost
        # more than the kernels of a small system: skip it while the objects and the scalars the
        # functional is built from are the ones of the previous call (a barostat changes rho0 / a).
        quick = (id(hamiltonian), id(config), getattr(config, "rho0", None), getattr(config, "a", None),
                 getattr(config, "kappa", None), getattr(config, "sigma", None), id(getattr(config, "chi", None)),
                 tuple(float(x) for x in (m if m is not None else (getattr(config, "m", None) or ()))),
                 getattr(config, "dielectric_const", None))
        if quick == getattr(self, "_interaction_quick", None):
            return
        A, c, m_cfg, sigma, conv = self._interaction(config, hamiltonian)
        if m is not None:
            m_cfg = np.asarray(list(m), dtype=np.float64)
        key = (A.tobytes(), c.tobytes(), m_cfg.tobytes(), sigma, conv)
        self._interaction_quick = quick
        # keep the probed objects alive: their id()s are part of the key and must not be recycled
        self._interaction_refs = (hamiltonian, config, getattr(config, "chi", None))
        if key == self._interaction_key:
            return
        dp = ctypes.POINTER(ctypes.c_double)
        Ac = np.ascontiguousarray(A, dtype=np.float64)
        _lib.check(self.lib.hymd_ctx_set_interaction(
            self._ctx, Ac.ctypes.data_as(dp), c.ctypes.data_as(dp), m_cfg.ctypes.data_as(dp),
            sigma, conv))
        self._interaction_key = key

    def set_box(self, box):
        self.BoxSize = np.asarray(box, dtype=np.float64).reshape(3).copy()
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.check(self.lib.hymd_ctx_set_box(self._ctx, self.BoxSize.ctypes.data_as(dp)))
        self._sort_key = None

    def secondary_pme(self, config):
        """A second particle-mesh context on the same mesh for PME calls on ANOTHER particle set with its own
        meshes -- the peptide backbone dipoles of ``main.py:1060-1095`` (``phi_dipoles``, ``psi_dipoles``, ... come
        from ``pm.create``).  The reference keeps those results in separate pmesh fields so that ``phi_q`` / ``psi``
        of the real charges survive for the energy print (``field.py:697-699``); here the second set gets its own
        one-type context (charge density, spectrum, potential, three field meshes), created on first use
        (collective on several GPUs) and re-boxed when a barostat has changed the box."""
        sec = getattr(self, "_secondary", None)
        if sec is None:
            from types import SimpleNamespace
            cfg = SimpleNamespace(
                n_types=1, mesh_size=[int(x) for x in self.Nmesh], box_size=self.BoxSize.copy(),
                dtype=np.float64 if self.np_dtype == np.dtype("f8") else np.float32,
                sigma=float(config.sigma), coulombtype="PIC_Spectral", m=[1.0],
                coulomb_constant=float(config.coulomb_constant), dielectric_const=float(config.dielectric_const))
            sec = ParticleMesh(self.Nmesh, BoxSize=self.BoxSize, dtype=self.np_dtype, comm=self.comm, config=cfg)
            self._secondary = sec
        if not np.array_equal(sec.BoxSize, self.BoxSize):
            sec.set_box(self.BoxSize)
        return sec

    def close(self):
        sec = getattr(self, "_secondary", None)
        if sec is not None:
            sec.close()
            self._secondary = None
        if self._ctx:
            self.lib.hymd_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- pmesh-like surface --------------------------------------------------------------
    def decompose(self, pos, smoothing=None):
        return Layout(self, 0 if pos is None else len(pos), smoothing)

    def create(self, kind="real", value=0.0):
        """``pm.create`` (``main.py:488-498``): a free-standing zero mesh (torch tensor)."""
        nxl = int(self.Nmesh[0]) // self.world_size
        if kind == "real":
            return torch.full((nxl, int(self.Nmesh[1]), int(self.Nmesh[2])), float(value),
                              dtype=self.dtype, device=self.device)
        return torch.full((int(self.Nmesh[0]), int(self.Nmesh[1]) // self.world_size,
                           int(self.Nmesh[2]) // 2 + 1), value, dtype=self.cdtype,
                          device=self.device)

    def field(self, field_id, t=0, d=0, kind="real"):
        return MeshField(self, field_id, t, d, kind)

    # ---- device plumbing -----------------------------------------------------------------
    @property
    def stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def as_device(self, x, dtype=None, shape=None, role=None):
        """C-contiguous device tensor of the mesh dtype (numpy in any memory order is copied).
        Large numpy / pageable CPU inputs go through a persistent pinned staging buffer per ``role``
        (a pageable cudaMemcpy runs at a fraction of the PCIe rate)."""
        dtype = dtype or self.dtype
        if isinstance(x, torch.Tensor):
            t = x
            if t.device != self.device:
                t = t.to(self.device, non_blocking=True)
            if t.dtype != dtype:
                t = t.to(dtype)
            t = t.contiguous()
        else:
            npd = {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32}[dtype]
            src = torch.from_numpy(np.ascontiguousarray(x, dtype=npd))
            if role is not None and src.numel() * src.element_size() >= (1 << 20):
                pin = self._pinned(role, src.shape, dtype)
                pin.copy_(src)
                t = pin.to(self.device, non_blocking=True)
                self._pin_events[role].record(torch.cuda.current_stream(self.device))
            else:
                t = src.to(self.device, non_blocking=True)
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def _pinned(self, role, shape, dtype):
        """Pinned host staging buffer for ``role`` (re-used from call to call; waits for the transfer
        that last used it)."""
        if not hasattr(self, "_pin_bufs"):
            self._pin_bufs, self._pin_events = {}, {}
        buf = self._pin_bufs.get(role)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(tuple(shape), dtype=dtype).pin_memory()
            self._pin_bufs[role] = buf
            self._pin_events[role] = torch.cuda.Event()
        else:
            self._pin_events[role].synchronize()
        return buf

    def to_host(self, dev_tensor, out, role):
        """Device result -> caller's numpy array in place, through a pinned staging buffer."""
        if dev_tensor.numel() * dev_tensor.element_size() < (1 << 20):
            out[...] = dev_tensor.cpu().numpy()
            return
        pin = self._pinned(role, dev_tensor.shape, dev_tensor.dtype)
        pin.copy_(dev_tensor, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        out[...] = pin.numpy()

    @staticmethod
    def _fingerprint(x):
        """Key under which ``sort`` recognises "the positions the bins were built from".

        torch tensors: storage pointer + torch's version counter + the library's own write epoch
        (``hymd_md_kick_drift`` / ``hymd_bonded_inner_step`` write through raw pointers, which torch
        does not see).  numpy arrays have no version counter: the key is the identity of the array
        plus a hash of a strided sample of <= 65536 rows (every row for small systems), so an array
        mutated in place between ``update_field`` and a later read-out is re-binned.  Contract for
        callers that change fewer rows than the sample stride: pass a new array or call
        ``pm.reset_order()``."""
        if isinstance(x, torch.Tensor):
            return ("t", x.data_ptr(), x._version, _lib.write_epoch(x), tuple(x.shape), x.dtype)
        a = np.asarray(x)
        n = a.shape[0]
        probe = 0
        if a.ndim == 2 and n:
            step = max(1, n // 65536)
            probe = hash(np.ascontiguousarray(a[::step]).tobytes())
        return ("n", id(x), a.ctypes.data, a.shape, a.dtype.str, probe)

    def sort(self, positions, types, charges=None, force=False, cycle=None):
        """Bin the local particles (all types at once).  ``update_field`` always re-bins
        (``force=True``: it is the call that opens a step, ``main.py:976-996``);
        ``compute_field_force`` / ``update_field_force_q`` called afterwards on the same
        positions object reuse its bins (``types=None`` means "whatever the bins hold").

        ``cycle`` (``update_field`` only: 0 / 1 = ``compute_potential``): the paint and the field cycle follow
        the binning inside the same library call (``hymd_update_cycle``), which small systems replay as one
        CUDA graph."""
        pk = self._fingerprint(positions)
        tk = None if types is None else self._fingerprint_types(types)
        if not force and pk == self._sort_key and (tk is None or tk == self._sort_types_key):
            if charges is not None and not self._sorted_with_charges:
                q = self.as_device(charges, shape=(self._n_local,))
                _lib.check(self.lib.hymd_set_charges(self._ctx, ctypes.c_void_p(q.data_ptr()),
                                                     self.stream))
                self._sorted_with_charges = True
                self._keep = (self._keep[0], self._keep[1], q)
            return
        pos = self.as_device(positions, role="positions")
        if pos.ndim != 2 or pos.shape[1] != 3:
            raise ValueError(f"positions must be (N,3), got {tuple(pos.shape)}")
        n = pos.shape[0]
        # Consecutive MD steps pass the same types object (HyMD reads the types once,
        # main.py:72-125): the library then re-bins starting from the previous cell order and
        # does not need the types again (they ride along in the sorted records).
        otk = ("none", n) if types is None else tk
        reuse = self._order_types_key is not None and otk == self._order_types_key \
            and n == self._n_local and (n > 0 or self.world_size > 1)
        ty = None
        if not reuse:
            if types is None:
                ty = torch.zeros(n, dtype=torch.int32, device=self.device)
            else:
                ty = self.as_device(types, dtype=torch.int32, shape=(n,))
        q = None if charges is None else self.as_device(charges, shape=(n,), role="charges")
        args = (self._ctx, ctypes.c_void_p(pos.data_ptr()),
                ctypes.c_void_p(ty.data_ptr()) if ty is not None else None,
                ctypes.c_void_p(q.data_ptr()) if q is not None else None, n,
                _lib.SORT_REUSE_ORDER if reuse else 0)
        if cycle is None:
            _lib.check(self.lib.hymd_sort_particles_ex(*args, self.stream))
        else:
            _lib.check(self.lib.hymd_update_cycle(*args, int(cycle), self.stream))
        self._keep = (pos, ty, q)   # inputs must outlive the asynchronous kernels
        self._sort_key, self._sort_types_key = pk, tk
        self._order_types_key = otk
        self._n_local = n
        self._sorted_with_charges = charges is not None

    def _fingerprint_types(self, types):
        if isinstance(types, torch.Tensor):
            return ("t", types.data_ptr(), types._version, tuple(types.shape))
        a = np.asarray(types)
        return ("n", id(types), a.ctypes.data, a.shape)

    def _materialize(self, want_phi=False, want_phi_fourier=False, want_v_ext=False, want_psi=False):
        if want_phi or want_phi_fourier or want_v_ext or want_psi:
            _lib.check(self.lib.hymd_materialize(self._ctx, int(want_phi), int(want_phi_fourier),
                                                 int(want_v_ext), int(want_psi), self.stream))

    def laplacian(self):
        """``phi_laplacian[t][d] = c2r(-k_d^2 phi_fourier[t])`` for the current spectra
        (``comp_laplacian``, ``field.py:406-425``); cached until the next paint."""
        _lib.check(self.lib.hymd_laplacian(self._ctx, self.stream))

    def field_pressure_sums(self, A, c, type_charges=None):
        """Local sums ``[sum V_t phi_t, sum V_t lap_t,x, .. y, .. z]`` over this slab's cells with
        ``V_t = c_t + sum_j A_tj phi_j (+ q_t psi)`` (field terms of ``pressure.py:105-127``)."""
        dp = ctypes.POINTER(ctypes.c_double)
        A = np.ascontiguousarray(A, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        q = None if type_charges is None else np.ascontiguousarray(type_charges, dtype=np.float64)
        out = (ctypes.c_double * 4)()
        _lib.check(self.lib.hymd_field_pressure(
            self._ctx, A.ctypes.data_as(dp), c.ctypes.data_as(dp),
            None if q is None else q.ctypes.data_as(dp), out, self.stream))
        return np.array(list(out), dtype=np.float64)

    def _view(self, field_id, t, d, kind):
        ptr = ctypes.c_void_p()
        dims = (ctypes.c_int64 * 3)()
        pitch = (ctypes.c_int64 * 3)()
        _lib.check(self.lib.hymd_get_field(self._ctx, field_id, t, d, ctypes.byref(ptr), dims, pitch))
        if kind == "complex":
            esz = 16 if self.np_dtype == np.dtype("f8") else 8
            typestr = "<c16" if esz == 16 else "<c8"
        else:
            esz = self.np_dtype.itemsize
            typestr = "<f8" if esz == 8 else "<f4"
        view = _CudaView(ptr.value, list(dims), [p * esz for p in pitch], typestr, self)
        return torch.as_tensor(view, device=self.device)

    def status(self):
        out = (ctypes.c_int64 * 4)()
        _lib.check(self.lib.hymd_ctx_status(self._ctx, out))
        return {"max_cell_count": out[0], "out_of_slab": out[1], "n_local": out[2],
                "potential_rows": out[3]}

    def paths(self):
        """Which kernels this context runs: fused x-line, plane transforms, slab pipeline,
        NVLink peer-memory exchange."""
        out = (ctypes.c_int32 * 4)()
        _lib.check(self.lib.hymd_ctx_paths(self._ctx, out))
        return {"xline": bool(out[0]), "plane": bool(out[1]), "slab": bool(out[2]), "p2p": bool(out[3])}

    def exchange_cost(self):
        """Guests sent by every rank in the last ``sort`` (length ``world_size``); zeros on one GPU."""
        P = max(self.world_size, 1)
        cost = np.zeros(P, dtype=np.int64)
        if P == 1 or self._ctx is None:
            return cost
        sent = (ctypes.c_int64 * P)()
        _lib.check(self.lib.hymd_exchange_cost(self._ctx, sent, self.stream))
        mine = torch.zeros(P, dtype=torch.int64)
        mine[self.rank] = int(sum(sent))
        return self.world.allreduce(mine).cpu().numpy().astype(np.int64)

    def set_timing(self, enable=True):
        _lib.check(self.lib.hymd_ctx_set_timing(self._ctx, int(bool(enable))))

    def timings(self):
        """{phase: (milliseconds, intervals)} accumulated since the previous call."""
        n = len(_lib.PHASES)
        ms = (ctypes.c_double * n)()
        calls = (ctypes.c_int64 * n)()
        _lib.check(self.lib.hymd_ctx_get_timings(self._ctx, ms, calls))
        return {_lib.PHASES[i]: (ms[i], calls[i]) for i in range(n) if calls[i]}

    def launch_count(self):
        return int(self.lib.hymd_launch_count(self._ctx))

    def set_graph(self, mode):
        """Graph replay of the per-step field update (``hymd_ctx_set_graph``): ``False`` / ``True`` / ``"auto"``."""
        m = -1 if mode == "auto" else int(bool(mode))
        _lib.check(self.lib.hymd_ctx_set_graph(self._ctx, m))

    def graph_stats(self):
        out = (ctypes.c_int64 * 4)()
        _lib.check(self.lib.hymd_ctx_graph_stats(self._ctx, out))
        return {"replayed": int(out[0]), "recorded": int(out[1]), "eager": int(out[2]), "alive": int(out[3])}

    def reset_order(self):
        """Forget the cached cell order / bins (the caller's particle order changed)."""
        self._sort_key = None
        self._order_types_key = None
        _lib.check(self.lib.hymd_ctx_reset_order(self._ctx))

    def _allreduce_scalar(self, s):
        return self.world.allreduce(s).item()

    def check(self):
        """Synchronize and raise if a device-side condition was flagged (guest capacity exceeded, a
        peer-memory barrier timed out); the per-step calls check the same flags without synchronizing."""
        _lib.check(self.lib.hymd_ctx_check(self._ctx))

    def migrate(self, positions, *arrays, routing_positions=None):
        """Re-home per-particle arrays on the rank owning their slab (``Layout.exchange`` of
        ``domain_decomposition``, ``field.py:1165-1178``).  All arrays have one row per local
        particle and are permuted identically; ``routing_positions`` (default: ``positions``)
        decides the destination.  Returns new arrays (torch CUDA tensors; numpy in -> numpy out)."""
        arrays = (positions,) + tuple(arrays)
        route = positions if routing_positions is None else routing_positions
        route_d = self.as_device(route)
        n = route_d.shape[0]
        n_new = ctypes.c_int64(0)
        _lib.check(self.lib.hymd_migrate_plan(self._ctx, ctypes.c_void_p(route_d.data_ptr()), n,
                                              ctypes.byref(n_new), self.stream))
        out = []
        for a in arrays:
            was_numpy = not isinstance(a, torch.Tensor)
            t = torch.as_tensor(np.ascontiguousarray(a)) if was_numpy else a
            t = t.to(self.device).contiguous()
            if t.shape[0] != n:
                raise ValueError(f"array with {t.shape[0]} rows in a migration of {n} particles")
            if t.dtype == torch.bool:
                t = t.to(torch.uint8)
            row_bytes = t.element_size() * int(np.prod(t.shape[1:], dtype=np.int64))
            o = torch.empty((n_new.value,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
            _lib.check(self.lib.hymd_migrate_apply(self._ctx, ctypes.c_void_p(t.data_ptr()),
                                                   ctypes.c_void_p(o.data_ptr()), row_bytes,
                                                   self.stream))
            if isinstance(a, torch.Tensor) and a.dtype == torch.bool:
                o = o.to(torch.bool)
            if was_numpy:
                o = o.cpu().numpy()
            elif a.device != self.device:
                o = o.to(a.device)
            out.append(o)
        self._sort_key = None
        self._order_types_key = None
        _lib.check(self.lib.hymd_ctx_reset_order(self._ctx))
        return tuple(out)


def _has_density_params(config):
    return getattr(config, "n_particles", None) is not None and \
        getattr(config, "box_size", None) is not None and getattr(config, "hamiltonian", None)

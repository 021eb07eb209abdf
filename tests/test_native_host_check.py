"""The device arithmetic of csrc/bonded.cuh and csrc/md.cuh, executed on the CPU.

The per-particle functions of the bonded / integrator / thermostat kernels are ``__host__ __device__``;
tests/native/host_check.cpp wraps exactly that source in plain loops and is compiled here with g++.
This checks, without a GPU, what the kernels compute per particle (term lists and slots, minimum image,
angle / dihedral algebra, cosine series, kick / drift / wrap rounding, CSVR scale) against the oracle
and the reference's golden vectors; the launch geometry and reductions are covered by the ``-m gpu``
tests (tests/test_zgpu_md.py)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import bonded_oracle as bo
from oracle import thermostat_oracle as to

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "bonded_golden.npz"))
TG = np.load(os.path.join(HERE, "golden", "thermostat_golden.npz"))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("native") / "libhost_check.so")
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "native", "host_check.cpp")], check=True)
    return ctypes.CDLL(so)


def i32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


def run_bonded(lib, kind, r, box, idx, par, dtype4=None, real=np.float64):
    r = np.ascontiguousarray(r, dtype=real)
    n, nt = r.shape[0], len(idx[0])
    idx = [i32(x) for x in idx] + [i32(np.zeros(nt))] * (4 - len(idx))
    par = np.ascontiguousarray(par, dtype=np.float64)
    dt = i32(dtype4 if dtype4 is not None else np.zeros(nt))
    f = np.zeros_like(r)
    out = np.zeros(4)
    box = np.ascontiguousarray(box, dtype=np.float64)
    vp = ctypes.c_void_p
    rc = lib.host_bonded(kind, int(real == np.float64), r.ctypes.data_as(vp), box.ctypes.data_as(vp),
                         ctypes.c_longlong(n), ctypes.c_longlong(nt), *[x.ctypes.data_as(vp) for x in idx],
                         par.ctypes.data_as(vp), dt.ctypes.data_as(vp), f.ctypes.data_as(vp),
                         out.ctypes.data_as(vp))
    assert rc == 0
    return f, out


# float32: the kernel rounds the force to float32 once (<= 6e-8 relative to its own magnitude); the
# position differences are float32 subtractions in the Fortran, the oracle and the kernel alike
@pytest.mark.parametrize("real,tol", [(np.float64, 1e-12), (np.float32, 1e-7)])
def test_bonds_angles_dihedrals_match_oracle(lib, real, tol):
    r, box = G["chains/r"], G["chains/box"]
    rr = r.astype(real)
    a, b, r0, k = (G["chains/b2_" + x] for x in ("a", "b", "r0", "k"))
    f, out = run_bonded(lib, 2, rr, box, [a, b], np.stack([r0, k], 1), real=real)
    fo_, e, pr = bo.compute_bond_forces(rr, box, a, b, r0, k)
    scale = np.abs(fo_).max()
    assert np.abs(f - fo_).max() <= tol * scale
    assert out[0] == pytest.approx(e, rel=max(tol, 1e-12))
    np.testing.assert_allclose(out[1:], pr, rtol=0, atol=max(tol, 1e-12) * np.abs(pr).max())
    if real == np.float64:      # and the reference's own plain function, through the golden file
        assert np.abs(f - G["chains/b2_force"]).max() <= 1e-10 * scale

    a, b, c, t0, k = (G["chains/b3_" + x] for x in ("a", "b", "c", "t0", "k"))
    f, out = run_bonded(lib, 3, rr, box, [a, b, c], np.stack([t0, k], 1), real=real)
    fo_, e, pr = bo.compute_angle_forces(rr, box, a, b, c, t0, k)
    scale = np.abs(fo_).max()
    assert np.abs(f - fo_).max() <= tol * scale
    assert out[0] == pytest.approx(e, rel=max(tol, 1e-11))
    if real == np.float64:
        assert np.abs(f - G["chains/b3_force"]).max() <= 1e-9 * scale

    r, box = G["dih/r"], G["dih/box"]
    rr = r.astype(real)
    a, b, c, d, coeff = (G["dih/" + x] for x in ("a", "b", "c", "d", "coeff"))
    dt = np.zeros(len(a), dtype=int)
    dt[::3] = 2
    coeff = coeff.copy()
    coeff[::3, 0, 0] = 0.3
    coeff[::3, 0, 1] = 40.0
    coeff[1::3, 2] = 0.5 * coeff[1::3, 0]        # coil series switched on for a third of the terms
    coeff[1::3, 3] = 0.25 + coeff[1::3, 1]
    f, out = run_bonded(lib, 4, rr, box, [a, b, c, d], coeff.reshape(len(a), -1), dt, real=real)
    fo_, e = bo.compute_dihedral_forces(rr, box, a, b, c, d, coeff, dt)
    scale = np.abs(fo_).max()
    assert np.abs(f - fo_).max() <= tol * scale
    assert out[0] == pytest.approx(e, rel=max(tol, 1e-11))
    assert np.all(out[1:] == 0)


@pytest.mark.parametrize("real,tol", [(np.float64, 1e-12), (np.float32, 2e-6)])
def test_cbt_dihedrals_dipoles_and_redistribution_match_oracle(lib, real, tol):
    """dtype-1 dihedrals through the device source (cbt_eval / dipole_term / redistribute_term of csrc/bonded.cuh):
    the propensity part comes from the dihedral kernel (dtype 1 is evaluated like dtype 0 there), the bending term,
    the dipoles with their transfer matrices and the redistribution from the three new functions -- together they
    equal the oracle's restatement of compute_dihedral_forces.f90 + dipole_reconstruction.f90 + force.py:855-880."""
    rng = np.random.default_rng(21)
    n = 40
    box = np.array([3.0, 3.5, 4.0])
    r = np.mod(np.cumsum(rng.normal(scale=0.33, size=(n, 3)), axis=0) + 1.5, box).astype(real)
    a = np.arange(n - 3, dtype=np.int32)
    nt = len(a)
    coeff = np.zeros((nt, 6, 5))
    coeff[:, 0] = rng.uniform(0.5, 2.0, size=(nt, 5))
    coeff[:, 1] = rng.uniform(-1, 1, size=(nt, 5))
    coeff[:, 4] = rng.uniform(20, 60, size=(nt, 5))
    coeff[:, 5] = rng.uniform(-1, 1, size=(nt, 5))
    dt = rng.choice([0, 1, 1, 2], size=nt).astype(np.int32)
    dt[-1] = 1
    coeff[dt == 2, 0, :2] = [0.4, 30.0]
    last = np.zeros(nt, dtype=np.int32)
    last[-1] = 1
    last[nt // 2] = 1
    idx = [a, a + 1, a + 2, a + 3]
    f_main, out = run_bonded(lib, 4, r, box, idx, coeff.reshape(-1), dt, real)
    vp = ctypes.c_void_p
    f_add = np.zeros((n, 3), dtype=real)
    e_cbt = ctypes.c_double(0.0)
    dip = np.zeros((nt, 4, 3), dtype=real)
    tm = np.zeros((nt, 6, 3, 3), dtype=real)
    fd = rng.normal(size=(nt, 4, 3)).astype(real)
    f_beads = np.zeros((n, 3), dtype=real)
    par = np.ascontiguousarray(coeff.reshape(-1))
    rc = lib.host_cbt(int(real == np.float64), r.ctypes.data_as(vp), box.ctypes.data_as(vp), ctypes.c_longlong(n),
                      ctypes.c_longlong(nt), *[i32(x).ctypes.data_as(vp) for x in idx], par.ctypes.data_as(vp),
                      dt.ctypes.data_as(vp), last.ctypes.data_as(vp), f_add.ctypes.data_as(vp), ctypes.byref(e_cbt),
                      dip.ctypes.data_as(vp), tm.ctypes.data_as(vp), fd.ctypes.data_as(vp), f_beads.ctypes.data_as(vp))
    assert rc == 0
    fo, eo, dipo, tmo = bo.compute_dihedral_forces(r, box, *idx, coeff, dt, last, dipole_flag=1, full=True)
    scale = np.abs(fo).max()
    assert np.abs(f_main.astype(np.float64) + f_add.astype(np.float64) - fo).max() < tol * scale
    assert out[0] + e_cbt.value == pytest.approx(eo, rel=max(tol, 1e-12))
    assert np.abs(dip.astype(np.float64) - dipo.astype(np.float64)).max() < (1e-12 if real == np.float64 else 1e-6)
    assert np.abs(tm.astype(np.float64) - tmo.astype(np.float64)).max() < tol * max(1.0, np.abs(tmo).max())
    fb = bo.dipole_forces_redistribution(n, fd, tm, *idx, dt, last)
    assert np.abs(f_beads.astype(np.float64) - fb).max() < tol * max(1.0, np.abs(fb).max())
    # rows / matrices of dihedrals that are not dtype 1, and the second pair where the dihedral is not the last one
    assert np.all(dip[dt != 1] == 0) and np.all(tm[dt != 1] == 0)
    assert np.all(dip[(dt == 1) & (last == 0), 2:] == 0)


def test_reference_kats_through_the_device_source(lib):
    """test/test_force.py known answers (see tests/test_oracle_bonded.py) through bonded.cuh."""
    r, box = G["dppc/r"], G["dppc/box"]
    a, b, r0, k = (G["dppc/b2_" + x] for x in ("a", "b", "r0", "k"))
    f, out = run_bonded(lib, 2, r, box, [a, b], np.stack([r0, k], 1))
    np.testing.assert_allclose(f, G["dppc/b2_force"], rtol=0, atol=1e-11)
    assert out[0] == pytest.approx(float(G["dppc/b2_energy"]), abs=1e-12)
    a, b, c, t0, k = (G["dppc/b3_" + x] for x in ("a", "b", "c", "t0", "k"))
    f, out = run_bonded(lib, 3, r, box, [a, b, c], np.stack([t0, k], 1))
    np.testing.assert_allclose(f, G["dppc/b3_force"], rtol=0, atol=1e-10)
    assert out[0] == pytest.approx(float(G["dppc/b3_energy"]), abs=1e-11)
    r, box = G["ala/r"], G["ala/box"]
    a, b, c, d, coeff, dt = (G["ala/" + x] for x in ("a", "b", "c", "d", "coeff", "dtype"))
    f, out = run_bonded(lib, 4, r, box, [a, b, c, d], coeff.reshape(len(a), -1), dt)
    np.testing.assert_allclose(f, -G["ala/term_force_plain"].sum(axis=0), rtol=0, atol=1e-10)
    assert out[0] == pytest.approx(float(G["ala/term_energy"].sum()), abs=1e-11)


def test_term_list_rejects_bad_indices(lib):
    r = np.zeros((4, 3))
    vp = ctypes.c_void_p
    for a, b in (([0, 5], [1, 2]), ([0, 1], [0, 2]), ([-1], [2])):
        a, b = i32(a), i32(b)
        z = i32(np.zeros(len(a)))
        par = np.zeros((len(a), 2))
        f, out = np.zeros_like(r), np.zeros(4)
        box = np.ones(3)
        rc = lib.host_bonded(2, 1, r.ctypes.data_as(vp), box.ctypes.data_as(vp), ctypes.c_longlong(4),
                             ctypes.c_longlong(len(a)), a.ctypes.data_as(vp), b.ctypes.data_as(vp),
                             z.ctypes.data_as(vp), z.ctypes.data_as(vp), par.ctypes.data_as(vp),
                             z.ctypes.data_as(vp), f.ctypes.data_as(vp), out.ctypes.data_as(vp))
        assert rc == -1


@pytest.mark.parametrize("real", [np.float64, np.float32])
def test_kick_drift_matches_numpy_expressions(lib, real):
    """main.py:803-837 evaluated by numpy in the array dtype (what the reference does) vs md.cuh."""
    rng = np.random.default_rng(5)
    n = 1000
    box = np.array([3.0, 4.0, 5.0])
    x = (rng.random((n, 3)) * box).astype(real)
    v = rng.normal(scale=2.0, size=(n, 3)).astype(real)
    fs = [rng.normal(scale=300.0, size=(n, 3)).astype(real) for _ in range(3)]
    mass, dt = 72.0, 0.03
    vp = ctypes.c_void_p
    fptr = (vp * 3)(*[f.ctypes.data for f in fs])
    bx = np.ascontiguousarray(box)
    d = ctypes.c_double
    # inner step: summed forces, then drift + wrap
    v1, x1 = v.copy(), x.copy()
    lib.host_kick_drift(int(real == np.float64), v1.ctypes.data_as(vp), x1.ctypes.data_as(vp), fptr, 3, 0,
                        d(mass), d(dt), d(dt), bx.ctypes.data_as(vp), ctypes.c_longlong(n))
    v_ref = v + real(0.5 * dt) * ((fs[0] + fs[1] + fs[2]) / real(mass))
    x_ref = np.mod(x + real(dt) * v_ref, box.astype(real)[None, :])
    np.testing.assert_array_equal(v1, v_ref)
    assert np.abs(x1 - x_ref).max() <= 4 * np.finfo(real).eps * box.max()
    assert (x1 >= 0).all() and (x1 < box.astype(real)).all()
    # outer step: one kick per force in turn
    v2 = v.copy()
    lib.host_kick_drift(int(real == np.float64), v2.ctypes.data_as(vp), None, fptr, 2, 1, d(mass), d(5 * dt),
                        d(0.0), None, ctypes.c_longlong(n))
    v_ref = v + real(0.5 * 5 * dt) * (fs[0] / real(mass))
    v_ref = v_ref + real(0.5 * 5 * dt) * (fs[1] / real(mass))
    np.testing.assert_array_equal(v2, v_ref)


def _groups(names, groups):
    out = np.full(len(names), -1, dtype=np.int32)
    for i, g in enumerate(groups):
        for t in g:
            out[names == np.bytes_(t)] = i
    return out


@pytest.mark.parametrize("remove", [False, True])
@pytest.mark.parametrize("case,groups", [("all", None), ("abcd", [["A"], ["B"], ["C"], ["D"]]),
                                         ("abc_d", [["A", "B", "C"], ["D"]])])
def test_csvr_matches_reference_golden(lib, case, groups, remove):
    pre = f"mws/{case}_{'com' if remove else 'nocom'}"
    mass, gas, T0, dt, inner, tau = TG["mws/params"]
    v = np.ascontiguousarray(TG["mws/velocities"].copy())
    names = TG["mws/names"]
    grp = np.zeros(len(names), dtype=np.int32) if groups is None else _groups(names, groups)
    work = np.zeros(1)
    vp, d = ctypes.c_void_p, ctypes.c_double
    c = float(np.exp(-(dt * inner) / tau))
    for g, (R, S) in enumerate(zip(TG[pre + "/gauss"], TG[pre + "/chi2"])):
        lib.host_csvr(1, v.ctypes.data_as(vp), grp.ctypes.data_as(vp), g, ctypes.c_longlong(len(v)), d(mass),
                      d(1.5 * gas * T0), d(c), d(R), d(S), int(remove), work.ctypes.data_as(vp))
    np.testing.assert_allclose(v, TG[pre + "/v"], rtol=1e-12, atol=1e-14)
    assert work[0] == pytest.approx(float(TG[pre + "/work"]), abs=1e-10)


@pytest.mark.parametrize("mesh", [(8, 8, 8), (10, 12, 8), (9, 8, 11), (6, 5, 7)])
def test_gpe_kspace_arithmetic_matches_numpy(lib, mesh):
    """csrc/gpe.cuh (the four k-space modes of hymd_gpe_cycle, with the Nyquist rule and the padded
    z pitch of the k layout) executed on the CPU between numpy's raw rfftn and irfftn, against the
    oracle's expressions `c2r(op(k) * r2c(x))` (oracle/gpe_oracle.py = hymd/field.py:1006-1111)."""
    from oracle import pm_oracle as pmo
    nx, ny, nz = mesh
    box = np.array([3.5, 4.0, 3.0])
    sigma = 0.5
    rng = np.random.default_rng(3)
    nzc = nz // 2 + 1
    nzcp = (nzc + 1) // 2 * 2
    M = nx * ny * nz
    # tables exactly as context.cu build_tables lays them out
    def kax(n, L, m):
        idx = np.arange(m)
        ni = np.where(idx < (n + 1) // 2, idx, idx - n)
        return 2.0 * np.pi * ni / L
    ks = [kax(nx, box[0], nx), kax(ny, box[1], ny), kax(nz, box[2], nzc)]
    tab = np.concatenate([np.exp(-0.5 * sigma ** 2 * k ** 2) for k in ks] + ks)
    k = pmo.kgrid(mesh, box)
    k2 = pmo.knorm2_zeromode1(k)
    H = np.exp(-0.5 * sigma ** 2 * (k[0] ** 2 + k[1] ** 2 + k[2] ** 2))

    def pack(spec):                      # (F,nx,ny,nzc) complex -> [f][x][y][Nzcp] interleaved reals
        F = spec.shape[0]
        buf = np.zeros((F, nx, ny, nzcp), dtype=np.complex128)
        buf[..., :nzc] = spec
        buf[..., nzc:] = 7.0 + 3.0j      # padding must never leak into the result
        return np.ascontiguousarray(buf).view(np.float64).ravel()

    def unpack(flat, F):
        return flat.view(np.complex128).reshape(F, nx, ny, nzcp)[..., :nzc]

    def run(x, F, coef, use_h, div_k2, sign, want_s, want_v):
        raw = np.stack([np.fft.rfftn(x[f]) for f in range(F)])            # unnormalised r2c
        inp = pack(raw)
        out_s = np.full(F * nx * ny * nzcp * 2, np.nan) if want_s else None
        out_v = np.full(3 * F * nx * ny * nzcp * 2, np.nan) if want_v else None
        vp = ctypes.c_void_p
        lib.host_gpe_kspace(inp.ctypes.data_as(vp), out_s.ctypes.data_as(vp) if want_s else None,
                            out_v.ctypes.data_as(vp) if want_v else None, tab.ctypes.data_as(vp), nx, ny, nz, F,
                            ctypes.c_double(coef), use_h, div_k2, ctypes.c_double(sign))
        c2r = lambda spec: np.fft.irfftn(spec, s=mesh, axes=(0, 1, 2)) * M                 # unnormalised c2r
        s_real = np.stack([c2r(a) for a in unpack(out_s, F)]) if want_s else None
        v_real = np.stack([c2r(a) for a in unpack(out_v, 3 * F)]) if want_v else None
        return s_real, v_real

    def close(a, b):
        assert np.abs(a - b).max() <= 1e-11 * max(np.abs(b).max(), 1e-300)

    x = rng.normal(size=(1,) + tuple(mesh))
    xf = pmo.r2c(x[0])
    # filter: c2r(H * r2c(x))   (field.py:1008-1010)
    s_real, _ = run(x, 1, 1.0 / M, 1, 0, 1.0, True, False)
    close(s_real[0], pmo.c2r(H * xf, mesh))
    # gradient of the dielectric: c2r(+i k_d r2c(x))   (field.py:1027-1032)
    _, v_real = run(x, 1, 1.0 / M, 0, 0, 1.0, False, True)
    for d in range(3):
        close(v_real[d], pmo.c2r(1j * k[d] * xf, mesh))
    # iteration field: c2r(-i k_d r2c(x) / k^2)   (field.py:1049-1052)
    _, v_real = run(x, 1, 1.0 / M, 0, 1, -1.0, False, True)
    for d in range(3):
        close(v_real[d], pmo.c2r(-1j * k[d] * xf / k2, mesh))
    # potential and field in one pass   (field.py:1070-1078)
    s_real, v_real = run(x, 1, 1.0 / M, 0, 1, -1.0, True, True)
    close(s_real[0], pmo.c2r(xf / k2, mesh))
    for d in range(3):
        close(v_real[d], pmo.c2r(-1j * k[d] * (xf / k2), mesh))
    # filtered -grad of T potentials, 3T outputs ordered [3t+d]   (field.py:1099-1108)
    T = 3
    xs = rng.normal(size=(T,) + tuple(mesh))
    _, v_real = run(xs, T, 1.0 / M, 1, 0, -1.0, False, True)
    for t in range(T):
        for d in range(3):
            close(v_real[3 * t + d], pmo.c2r(-1j * k[d] * (H * pmo.r2c(xs[t])), mesh))


def test_gpe_cycle_plan_matches_reference_golden(lib):
    """The data flow of csrc/gpe.cu::gpe_cycle_t transcribed step by step -- raw (unnormalised) forward
    transforms, the gpe.cuh k-space kernel with exactly the (coef, use_h, div_k2, sign) arguments the
    driver passes, unnormalised inverse transforms, the pointwise updates with the reference's masks --
    reproduces the reference's own update_field_force_q_GPE (tests/golden/gpe_golden.npz).  This pins
    the plan (normalisation, signs, scalings, buffer roles); the CUDA glue itself is GPU-only."""
    from oracle import pm_oracle as pmo
    GG = np.load(os.path.join(HERE, "golden", "gpe_golden.npz"))
    pre = "gpe/gpe3"
    mesh, box = (10, 12, 8), np.array([3.5, 4.0, 3.0])
    sigma, k_e = 0.5, 138.935458
    eps_t, q_t, w, crit = np.array([5.0, 10.0, 80.0]), np.array([1.0, -1.0, 0.0]), 0.6, 1e-6
    nx, ny, nz = mesh
    nzc = nz // 2 + 1
    nzcp = (nzc + 1) // 2 * 2
    M = nx * ny * nz
    dv = np.prod(box) / M

    def kax(n, L, m):
        idx = np.arange(m)
        return 2.0 * np.pi * np.where(idx < (n + 1) // 2, idx, idx - n) / L
    ks = [kax(nx, box[0], nx), kax(ny, box[1], ny), kax(nz, box[2], nzc)]
    tab = np.concatenate([np.exp(-0.5 * sigma ** 2 * k ** 2) for k in ks] + ks)
    vp = ctypes.c_void_p

    def fwd(x):                                   # fft_forward: raw spectra in the padded k layout
        x = np.asarray(x).reshape((-1,) + mesh)
        buf = np.zeros((x.shape[0], nx, ny, nzcp), dtype=np.complex128)
        buf[..., :nzc] = np.fft.rfftn(x, axes=(1, 2, 3))
        return buf

    def inv(buf):                                 # fft_inverse: unnormalised c2r
        return np.fft.irfftn(buf[..., :nzc], s=mesh, axes=(1, 2, 3)) * M

    def kspace(inp, want_s, want_v, coef, use_h, div_k2, sign):
        F = inp.shape[0]
        flat = np.ascontiguousarray(inp).view(np.float64).ravel()
        out_s = np.zeros(F * nx * ny * nzcp * 2) if want_s else None
        out_v = np.zeros(3 * F * nx * ny * nzcp * 2) if want_v else None
        lib.host_gpe_kspace(flat.ctypes.data_as(vp), out_s.ctypes.data_as(vp) if want_s else None,
                            out_v.ctypes.data_as(vp) if want_v else None, tab.ctypes.data_as(vp), nx, ny, nz, F,
                            ctypes.c_double(coef), int(use_h), int(div_k2), ctypes.c_double(sign))
        s = out_s.view(np.complex128).reshape(F, nx, ny, nzcp) if want_s else None
        v = out_v.view(np.complex128).reshape(3 * F, nx, ny, nzcp) if want_v else None
        return s, v

    phi = GG[pre + "/phi"]
    pos, types_, charges = GG[pre + "/pos"], GG[pre + "/types"], GG[pre + "/charges"]
    T = len(eps_t)
    # --- the driver, line by line
    phi_q = pmo.cic_paint(pos, charges, mesh, box, np.float64, use_c=False) / dv          # paint_charges
    kS, _ = kspace(fwd(phi_q), True, False, 1.0 / M, True, False, 1.0)
    phi_q = inv(kS)[0]
    den = phi.sum(axis=0)
    eps = np.zeros(mesh)
    np.divide((eps_t[:, None, None, None] * phi).sum(axis=0), den, where=den > 1e-6, out=eps)   # gpe_eps_kernel
    _, kB = kspace(fwd(eps), False, True, 1.0 / M, False, False, 1.0)
    eta = inv(kB)
    mask = eps > 1e-6                                                                       # gpe_divide_kernel
    phi_q = np.where(mask, phi_q / np.where(mask, eps, 1.0), phi_q)
    eta = np.where(mask, eta / np.where(mask, eps, 1.0), eta)
    pol = np.zeros(mesh)
    it, delta = 0, 1.0
    while it < 100 and delta > crit:
        _, kB = kspace(fwd(phi_q + pol), False, True, 1.0 / M, False, True, -1.0)
        E = inv(kB)
        new = w * (-(eta * E).sum(axis=0)) + (1.0 - w) * pol                               # gpe_pol_kernel
        delta = np.abs(new - pol).max()
        pol = new
        it += 1
    eps0_inv = k_e * 4 * np.pi
    kS, kB = kspace(fwd(eps0_inv * (phi_q + pol)), True, True, 1.0 / M, False, True, -1.0)
    psi, E = inv(kS)[0], inv(kB)
    dot = (E ** 2).sum(axis=0)                                                              # gpe_vbar_kernel
    contrib = np.zeros(mesh)
    np.divide(dot, den, where=den > 1e-6, out=contrib)
    vbar = np.stack([q_t[t] * psi - (0.5 / eps0_inv) * (eps_t[t] - eps) * contrib for t in range(T)])
    _, kB = kspace(fwd(vbar), False, True, 1.0 / M, True, False, -1.0)
    fmesh = inv(kB)
    forces = np.zeros((len(pos), 3))
    for t in range(T):
        ind = types_ == t
        for d in range(3):
            forces[ind, d] = pmo.cic_readout(np.ascontiguousarray(fmesh[3 * t + d]), pos[ind], box, use_c=False)
    energy = dv * 0.5 / eps0_inv * np.sum(eps * dot)                                        # gpe_energy

    def close(a, b, tol=1e-9):
        assert np.abs(a - b).max() <= tol * np.abs(b).max()
    close(eps, GG[pre + "/phi_eps"], 1e-12)
    close(psi, GG[pre + "/psi"])
    close(dot, GG[pre + "/elec_dot"])
    close(vbar, GG[pre + "/Vbar_elec"])
    close(forces, GG[pre + "/elec_forces"])
    assert energy == pytest.approx(float(GG[pre + "/energy"]), rel=1e-9)


def test_f32_math_accuracy(lib):
    """csrc/bonded_f32.cuh (opt-in single-precision arithmetic for the fp32 build) against the float64
    oracle on the same float32 positions: bonds and angles within 1e-5 of the largest force (the
    north_star tolerance of the fp32 build), including chains that are straight to 1e-3 rad with
    180-degree equilibrium angles, where acos-based float arithmetic loses every digit."""
    rng = np.random.default_rng(17)
    box = np.array([6.0, 7.0, 5.0])
    vp = ctypes.c_void_p

    def run(r, a2, a3, t0v):
        n = len(r)
        r0, k2 = np.full(len(a2), 0.47), np.full(len(a2), 1250.0)
        k3 = np.full(len(a3), 25.0)
        h = vp()
        empty = i32(np.zeros(0))
        ia2, ib2 = i32(a2), i32(a2 + 1)
        ia3, ib3, ic3 = i32(a3), i32(a3 + 1), i32(a3 + 2)
        lib.hymd_bonded_create.argtypes = None
        rc = lib.hymd_bonded_create(ctypes.c_int64(n), ctypes.c_int64(len(a2)), ia2.ctypes.data_as(vp),
                                    ib2.ctypes.data_as(vp), r0.ctypes.data_as(vp), k2.ctypes.data_as(vp),
                                    ctypes.c_int64(len(a3)), ia3.ctypes.data_as(vp), ib3.ctypes.data_as(vp),
                                    ic3.ctypes.data_as(vp), t0v.ctypes.data_as(vp), k3.ctypes.data_as(vp),
                                    ctypes.c_int64(0), empty.ctypes.data_as(vp), empty.ctypes.data_as(vp),
                                    empty.ctypes.data_as(vp), empty.ctypes.data_as(vp), np.zeros(1).ctypes.data_as(vp),
                                    empty.ctypes.data_as(vp), ctypes.byref(h))
        assert rc == 0
        x = np.ascontiguousarray(r, dtype=np.float32)
        fb, fa = np.zeros_like(x), np.zeros_like(x)
        fptr = (vp * 3)(fb.ctypes.data, fa.ctypes.data, None)
        out = np.zeros(12)
        bx = np.ascontiguousarray(box)
        lib.host_inner_step_f32(h, x.ctypes.data_as(vp), None, None, bx.ctypes.data_as(vp), ctypes.c_double(72.0),
                                ctypes.c_double(0.0), 0, ctypes.c_double(0.0), fptr, out.ctypes.data_as(vp))
        lib.hymd_bonded_destroy(h)
        fbo, eb, prb = bo.compute_bond_forces(x, box, a2, a2 + 1, r0, k2)
        fao, ea, pra = bo.compute_angle_forces(x, box, a3, a3 + 1, a3 + 2, t0v, k3)
        return (fb, fbo, out[0], eb, out[1:4], prb), (fa, fao, out[4], ea, out[5:8], pra)

    # random-walk chains (all angles), 120 / 180 degree equilibria
    n_ch, L = 200, 10
    steps = rng.normal(size=(n_ch, L - 1, 3))
    steps *= 0.47 * (1.0 + 0.1 * rng.normal(size=(n_ch, L - 1, 1))) / np.linalg.norm(steps, axis=2, keepdims=True)
    r = (rng.random((n_ch, 1, 3)) * box + np.concatenate([np.zeros((n_ch, 1, 3)), np.cumsum(steps, 1)], 1)).reshape(-1, 3)
    r = np.mod(r, box)
    first = (np.arange(n_ch) * L)[:, None]
    a2 = (first + np.arange(L - 1)[None, :]).ravel()
    a3 = (first + np.arange(L - 2)[None, :]).ravel()
    t0v = np.radians(rng.choice([120.0, 180.0], size=len(a3)))
    for got, want, e, e_ref, pr, pr_ref in run(r, a2, a3, t0v):
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
        assert e == pytest.approx(e_ref, rel=1e-5)
        assert np.abs(pr - pr_ref).max() <= 1e-5 * max(np.abs(pr_ref).max(), np.abs(want).max())
    # nearly straight chains (bending 1e-3 .. 3e-2 rad), 180-degree equilibrium: the hard case in float
    base = rng.normal(size=(n_ch, 1, 3))
    base /= np.linalg.norm(base, axis=2, keepdims=True)
    wob = rng.normal(size=(n_ch, L - 1, 3)) * rng.choice([1e-3, 1e-2, 3e-2], size=(n_ch, 1, 1))
    steps = base + wob
    steps *= 0.47 * (1.0 + 0.05 * rng.normal(size=(n_ch, L - 1, 1))) / np.linalg.norm(steps, axis=2, keepdims=True)
    r = (rng.random((n_ch, 1, 3)) * box + np.concatenate([np.zeros((n_ch, 1, 3)), np.cumsum(steps, 1)], 1)).reshape(-1, 3)
    r = np.mod(r, box)
    (fb, fbo, *_), (fa, fao, e, e_ref, pr, pr_ref) = run(r, a2, a3, np.full(len(a3), np.pi))
    scale = max(np.abs(fbo).max(), np.abs(fao).max())
    assert np.abs(fa - fao).max() <= 1e-5 * scale
    assert np.abs(fb - fbo).max() <= 1e-5 * scale
    assert e == pytest.approx(e_ref, rel=1e-3)      # tiny energies (theta - pi)^2 of float32 positions

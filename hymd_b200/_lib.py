"""ctypes binding of libhymd_b200.so (C ABI in include/hymd_b200.h).

There is deliberately no fallback: if the CUDA library is missing or does not load, every
entry point of hymd_b200.field fails with an explicit error.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libhymd_b200.so")

HYMD_MAX_TYPES = 32
NCCL_ID_BYTES = 128

F32, F64 = 0, 1
SORT_REUSE_ORDER = 1
(FIELD_PHI, FIELD_PHI_FOURIER, FIELD_FORCE_MESH, FIELD_V_EXT, FIELD_PHI_Q, FIELD_PHI_Q_FOURIER,
 FIELD_PSI, FIELD_ELEC_FIELD, FIELD_PHI_LAPLACIAN, FIELD_GPE_EPS, FIELD_GPE_ELEC_DOT,
 FIELD_GPE_VBAR) = range(12)

# every symbol include/hymd_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "hymd_last_error", "hymd_abi_version", "hymd_nccl_unique_id", "hymd_ctx_create",
    "hymd_ctx_destroy", "hymd_ctx_set_box", "hymd_ctx_set_interaction", "hymd_sort_particles",
    "hymd_set_charges", "hymd_paint", "hymd_field_cycle", "hymd_readout", "hymd_pme_cycle",
    "hymd_materialize", "hymd_field_energy", "hymd_get_field", "hymd_ctx_status",
    "hymd_launch_count", "hymd_migrate_plan", "hymd_migrate_apply", "hymd_ctx_set_timing", "hymd_ctx_get_timings",
    "hymd_sort_particles_ex", "hymd_ctx_reset_order", "hymd_ctx_paths", "hymd_laplacian",
    "hymd_field_pressure",
    "hymd_bonded_create", "hymd_bonded_destroy", "hymd_bonded_forces", "hymd_bonded_launch_count", "hymd_bonded_inner_step", "hymd_bonded_set_cta", "hymd_bonded_set_math",
    "hymd_md_kick_drift", "hymd_velocity_moments", "hymd_velocity_moments_scratch_doubles",
    "hymd_csvr_apply", "hymd_cancel_com",
    "hymd_gpe_cycle", "hymd_gpe_energy",
    "hymd_local_group_id", "hymd_ctx_check", "hymd_exchange_cost",
    "hymd_bonded_set_last", "hymd_bonded_dipoles", "hymd_dipole_redistribute",
    "hymd_update_cycle", "hymd_ctx_set_graph", "hymd_ctx_graph_stats",
]
PHASES = ["sort", "paint", "fft_fwd", "kspace", "fft_inv", "ghost", "readout", "pme_paint",
          "pme_fft", "pme_kspace", "pme_readout", "alltoall", "halo", "migrate", "byproducts",
          "reserved"]


class HymdConfig(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("mesh", ctypes.c_int32 * 3),
        ("box", ctypes.c_double * 3),
        ("n_types", ctypes.c_int32),
        ("world_size", ctypes.c_int32),
        ("rank", ctypes.c_int32),
        ("pme", ctypes.c_int32),
        ("sigma", ctypes.c_double),
        ("elec_conversion", ctypes.c_double),
        ("A", ctypes.c_double * (HYMD_MAX_TYPES * HYMD_MAX_TYPES)),
        ("c", ctypes.c_double * HYMD_MAX_TYPES),
        ("m", ctypes.c_double * HYMD_MAX_TYPES),
    ]


class HymdGpeParams(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_int32),
        ("convergence_type", ctypes.c_int32),
        ("max_iter", ctypes.c_int32),
        ("pad", ctypes.c_int32),
        ("pol_mixing", ctypes.c_double),
        ("conv_crit", ctypes.c_double),
        ("coulomb_constant", ctypes.c_double),
        ("dielectric_type", ctypes.c_double * HYMD_MAX_TYPES),
        ("type_charges", ctypes.c_double * HYMD_MAX_TYPES),
    ]


class HymdError(RuntimeError):
    pass


# Tensors the library writes through raw pointers (hymd_md_kick_drift, hymd_bonded_inner_step) keep
# their torch ``_version``; every such write is recorded here so that the bin cache of
# ParticleMesh.sort (keyed on data pointer + version) sees it.
_write_epoch = {}


def mark_written(tensor):
    """Record that the library modified ``tensor``'s storage behind torch's back."""
    key = tensor.data_ptr()
    if len(_write_epoch) > 4096:
        _write_epoch.clear()
    _write_epoch[key] = _write_epoch.get(key, 0) + 1


def write_epoch(tensor):
    return _write_epoch.get(tensor.data_ptr(), 0)


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HymdError(
            f"{LIB_PATH} is missing: build it with `python -m hymd_b200.build` "
            "(hymd_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    P = ctypes.POINTER
    lib.hymd_last_error.restype = ctypes.c_char_p
    lib.hymd_last_error.argtypes = []
    lib.hymd_abi_version.restype = ctypes.c_int
    lib.hymd_nccl_unique_id.argtypes = [P(ctypes.c_uint8)]
    lib.hymd_local_group_id.argtypes = [ctypes.c_int, P(ctypes.c_uint8)]
    lib.hymd_ctx_check.argtypes = [vp]
    lib.hymd_ctx_create.argtypes = [P(HymdConfig), P(ctypes.c_uint8), P(vp)]
    lib.hymd_ctx_destroy.argtypes = [vp]
    lib.hymd_ctx_set_box.argtypes = [vp, P(dbl)]
    lib.hymd_ctx_set_interaction.argtypes = [vp, P(dbl), P(dbl), P(dbl), dbl, dbl]
    lib.hymd_sort_particles.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.hymd_set_charges.argtypes = [vp, vp, vp]
    lib.hymd_sort_particles_ex.argtypes = [vp, vp, vp, vp, i64, ctypes.c_int, vp]
    lib.hymd_ctx_reset_order.argtypes = [vp]
    lib.hymd_paint.argtypes = [vp, vp]
    lib.hymd_field_cycle.argtypes = [vp, ctypes.c_int, vp]
    lib.hymd_readout.argtypes = [vp, vp, vp]
    lib.hymd_update_cycle.argtypes = [vp, vp, vp, vp, i64, ctypes.c_int, ctypes.c_int, vp]
    lib.hymd_ctx_set_graph.argtypes = [vp, ctypes.c_int]
    lib.hymd_ctx_graph_stats.argtypes = [vp, P(i64)]
    lib.hymd_pme_cycle.argtypes = [vp, vp, ctypes.c_int, vp]
    lib.hymd_materialize.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
    lib.hymd_field_energy.argtypes = [vp, P(dbl), dbl, dbl, dbl, P(dbl), vp]
    lib.hymd_laplacian.argtypes = [vp, vp]
    lib.hymd_field_pressure.argtypes = [vp, P(dbl), P(dbl), P(dbl), P(dbl), vp]
    lib.hymd_get_field.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, P(vp), P(i64), P(i64)]
    lib.hymd_ctx_status.argtypes = [vp, P(i64)]
    lib.hymd_ctx_paths.argtypes = [vp, P(i32)]
    lib.hymd_launch_count.argtypes = [vp]
    lib.hymd_launch_count.restype = i64
    lib.hymd_ctx_set_timing.argtypes = [vp, ctypes.c_int]
    lib.hymd_ctx_get_timings.argtypes = [vp, P(dbl), P(i64)]
    lib.hymd_migrate_plan.argtypes = [vp, vp, i64, P(i64), vp]
    lib.hymd_migrate_apply.argtypes = [vp, vp, vp, i32, vp]
    lib.hymd_exchange_cost.argtypes = [vp, P(i64), vp]
    I32P, F64P = P(i32), P(dbl)
    lib.hymd_bonded_create.argtypes = [i64, i64, I32P, I32P, F64P, F64P, i64, I32P, I32P, I32P, F64P, F64P,
                                       i64, I32P, I32P, I32P, I32P, F64P, I32P, P(vp)]
    lib.hymd_bonded_destroy.argtypes = [vp]
    lib.hymd_bonded_forces.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, P(dbl), vp, F64P, vp]
    lib.hymd_bonded_inner_step.argtypes = [vp, ctypes.c_int, vp, vp, vp, P(dbl), dbl, dbl, ctypes.c_int, dbl,
                                           P(vp), F64P, vp]
    lib.hymd_bonded_set_cta.argtypes = [vp, ctypes.c_int]
    lib.hymd_bonded_set_last.argtypes = [vp, I32P]
    lib.hymd_bonded_dipoles.argtypes = [vp, ctypes.c_int, vp, P(dbl), vp, vp, vp]
    lib.hymd_dipole_redistribute.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp]
    lib.hymd_bonded_set_math.argtypes = [vp, ctypes.c_int]
    lib.hymd_bonded_launch_count.argtypes = [vp]
    lib.hymd_bonded_launch_count.restype = i64
    lib.hymd_md_kick_drift.argtypes = [ctypes.c_int, vp, vp, P(vp), ctypes.c_int, ctypes.c_int, dbl, dbl, dbl,
                                       P(dbl), i64, vp]
    lib.hymd_velocity_moments.argtypes = [ctypes.c_int, vp, I32P, ctypes.c_int, i64, F64P, F64P, vp]
    lib.hymd_velocity_moments_scratch_doubles.argtypes = []
    lib.hymd_velocity_moments_scratch_doubles.restype = i64
    lib.hymd_csvr_apply.argtypes = [ctypes.c_int, vp, I32P, ctypes.c_int, i64, F64P, dbl, dbl, dbl, dbl, dbl,
                                    ctypes.c_int, F64P, vp]
    lib.hymd_cancel_com.argtypes = [ctypes.c_int, vp, i64, F64P, dbl, vp]
    lib.hymd_gpe_cycle.argtypes = [vp, P(HymdGpeParams), vp, P(i32), vp]
    lib.hymd_gpe_energy.argtypes = [vp, dbl, P(dbl), vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("hymd_last_error", "hymd_launch_count", "hymd_bonded_launch_count", "hymd_bonded_inner_step", "hymd_bonded_set_cta", "hymd_bonded_set_math",
                        "hymd_velocity_moments_scratch_doubles"):
            fn.restype = ctypes.c_int
    if lib.hymd_abi_version() != 1:
        raise HymdError(f"libhymd_b200.so ABI {lib.hymd_abi_version()} != 1")
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        msg = load().hymd_last_error().decode("utf-8", "replace")
        raise HymdError(f"libhymd_b200 error {status}: {msg}")

"""hymd_b200 — B200-native particle-mesh field-force cycle for HyMD.

Drop-in for ``hymd/field.py`` (``update_field`` / ``compute_field_force`` /
``update_field_force_q`` and friends) backed by hand-written sm_100a kernels
behind a C ABI (``include/hymd_b200.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"

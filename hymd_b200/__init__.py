"""hymd_b200 — B200-native particle-mesh field-force cycle for HyMD.

Drop-in for ``hymd/field.py`` (``update_field`` / ``compute_field_force`` /
``update_field_force_q`` and friends) backed by hand-written sm_100a kernels
behind a C ABI (``include/hymd_b200.h``).  There is no CPU fallback.

=================  ==========================================================================
module             mirrors (reference, HyMD v2.2.0)
=================  ==========================================================================
``field``          ``hymd/field.py``: initialize_pm, update_field, compute_field_force,
                   update_field_force_q, compute_field_and_kinetic_energy, comp_laplacian,
                   domain_decomposition
``pm``             the pmesh objects main.py threads through them (ParticleMesh, Layout, fields)
``pressure``       ``hymd/pressure.py``: comp_pressure
``barostat``       ``hymd/barostat.py`` / ``barostat_scr.py``: isotropic, semiisotropic
``force``          the f2py kernels cbf / caf / cdf: compute_{bond,angle,dihedral}_forces
``thermostat``     ``hymd/thermostat.py``: csvr_thermostat, cancel_com_momentum, generate_initial_velocities
``md``             ``hymd/integrator.py`` and the rRESPA loop of ``main.py:801-1169`` (RespaMD)
``hamiltonian``    ``hymd/hamiltonian.py``; ``config`` the field-relevant slice of ``input_parser.Config``
``file_io``        ``hymd/file_io.py``: OutDataset, store_static, store_data, distribute_input (needs h5py)
=================  ==========================================================================
"""
__version__ = "0.1.0"

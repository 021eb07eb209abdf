/* hymd_b200.h — C ABI of the B200-native particle-mesh field-force cycle.
 *
 * This is the drop-in boundary for the hot path of HyMD's hymd/field.py.  The
 * reference has no FFI of its own for this path (it calls the third-party
 * pmesh/PFFT Python packages); each entry point below therefore cites the
 * reference Python call sites it replaces (paths relative to the HyMD repo).
 *
 * Conventions
 *  - every function returns 0 on success or a negative hymd_status code; the
 *    message for the last failure on the calling thread is hymd_last_error().
 *  - all pointers named d_* are DEVICE pointers on the context's GPU; plain
 *    pointers are host memory.  No ownership is transferred.  Field buffers are
 *    owned by the context and exposed through hymd_get_field().
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Calls are asynchronous with respect to the host unless stated otherwise.
 *  - one context per GPU / rank; a context is not thread-safe.
 *  - real fields are C-ordered (x slowest, z fastest).  With world_size > 1 the
 *    mesh is split into contiguous slabs along x (the slowest storage axis):
 *    rank r owns planes [r*Nx/P, (r+1)*Nx/P) and the particles inside it.
 */
#ifndef HYMD_B200_H
#define HYMD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYMD_B200_ABI_VERSION 1
#define HYMD_MAX_TYPES 32
#define HYMD_NCCL_UNIQUE_ID_BYTES 128

typedef enum {
    HYMD_OK = 0,
    HYMD_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
    HYMD_ERR_CUDA = -2,      /* CUDA runtime or cuFFT failure */
    HYMD_ERR_NCCL = -3,      /* NCCL failure or libnccl not loadable */
    HYMD_ERR_STATE = -4,     /* call sequence violated (e.g. readout before sort) */
    HYMD_ERR_CAPACITY = -5,  /* particle outside the local slab / capacity overflow */
    HYMD_ERR_NOMEM = -6
} hymd_status;

typedef enum { HYMD_F32 = 0, HYMD_F64 = 1 } hymd_dtype;

/* Which context-owned buffer hymd_get_field() returns.  Shapes are for the local slab
 * (nxl = Nx / world_size planes); "pad" layouts are described in DESIGN.md. */
typedef enum {
    HYMD_FIELD_PHI = 0,          /* [t]      real  (nxl,Ny,Nz): raw painted density / dV after
                                    hymd_paint; FILTERED density after hymd_materialize */
    HYMD_FIELD_PHI_FOURIER = 1,  /* [t]      cplx  k-layout: H * r2c(phi)/M   (materialized) */
    HYMD_FIELD_FORCE_MESH = 2,   /* [t][d]   real  ghost-padded (nxl+1,Ny+1,Nzp) */
    HYMD_FIELD_V_EXT = 3,        /* [t]      real  (nxl,Ny,Nz)              (materialized) */
    HYMD_FIELD_PHI_Q = 4,        /*          real  (nxl,Ny,Nz) unfiltered charge density */
    HYMD_FIELD_PHI_Q_FOURIER = 5,/*          cplx  k-layout: H * r2c(phi_q)/M */
    HYMD_FIELD_PSI = 6,          /*          real  (nxl,Ny,Nz) electrostatic potential */
    HYMD_FIELD_ELEC_FIELD = 7,   /* [d]      real  ghost-padded (nxl+1,Ny+1,Nzp) */
    HYMD_FIELD_PHI_LAPLACIAN = 8,/* [t][d]   real  (nxl,Ny,Nz): c2r(-k_d^2 phi_fourier[t]) (hymd_laplacian) */
    HYMD_FIELD_GPE_EPS = 9,      /*          real  (nxl,Ny,Nz) relative dielectric phi_eps   (hymd_gpe_cycle) */
    HYMD_FIELD_GPE_ELEC_DOT = 10,/*          real  (nxl,Ny,Nz) |E|^2                         (hymd_gpe_cycle) */
    HYMD_FIELD_GPE_VBAR = 11     /* [t]      real  (nxl,Ny,Nz) Vbar_elec[t]                  (hymd_gpe_cycle) */
} hymd_field_id;

typedef struct {
    int32_t struct_size;        /* = sizeof(hymd_config), for ABI checking */
    int32_t dtype;              /* hymd_dtype: config.dtype, main.py:56-66 / field.py:32-35 */
    int32_t mesh[3];            /* config.mesh_size via np.full(3, mesh_size), field.py:571 */
    double box[3];              /* config.box_size [nm] */
    int32_t n_types;            /* config.n_types */
    int32_t world_size;         /* number of slabs (GPUs); 1 = single GPU */
    int32_t rank;               /* this slab */
    int32_t pme;                /* 1: allocate the PME buffers (config.coulombtype == "PIC_Spectral") */
    double sigma;               /* config.sigma: H(k) = exp(-sigma^2 k^2/2), hamiltonian.py:57-66 */
    double elec_conversion;     /* coulomb_constant / dielectric_const, field.py:360 */
    /* Affine external potential  V_t = sum_j A[t][j]*phi~_j + c[t]  (hamiltonian.py:402-412,
     * 470-473: A_tj = 1/(kappa rho0) + chi_tj/rho0, c_t = -a/(kappa rho0)).  Row-major T x T. */
    double A[HYMD_MAX_TYPES * HYMD_MAX_TYPES];
    double c[HYMD_MAX_TYPES];
    double m[HYMD_MAX_TYPES];   /* config.m: per-type paint mass, field.py:574 */
} hymd_config;

typedef struct hymd_ctx hymd_ctx;

/* Message describing the last error on this thread ("" if none). */
const char* hymd_last_error(void);
int hymd_abi_version(void);

/* Rank 0 obtains an id, the host program broadcasts it (any transport), every rank passes it
 * to hymd_ctx_create.  Replaces the MPI communicator handed to pmesh (field.py:45-47). */
int hymd_nccl_unique_id(uint8_t id[HYMD_NCCL_UNIQUE_ID_BYTES]);

/* "Virtual slabs": the world_size ranks are world_size host THREADS of this process (each with its own
 * context and stream, on one GPU or several) instead of one process per GPU.  Writes an id that
 * hymd_ctx_create accepts in place of an NCCL id; every thread then passes it with its own rank.  Peer
 * memory is plain device memory of the same address space and barriers synchronise the stream and meet
 * on the host, so the whole sharded pipeline (transposes, halos, per-step particle routing) runs -- and
 * is tested -- on a single-GPU box.  Nothing in the reference corresponds (it is always one MPI rank
 * per process). */
int hymd_local_group_id(int world_size, uint8_t id[HYMD_NCCL_UNIQUE_ID_BYTES]);

/* initialize_pm (field.py:10-149): builds cuFFT plans and every mesh buffer for this slab.
 * nccl_id may be NULL iff world_size == 1.  Synchronous. */
int hymd_ctx_create(const hymd_config* cfg, const uint8_t* nccl_id, hymd_ctx** out);
int hymd_ctx_destroy(hymd_ctx* ctx);

/* Barostat box rescale (barostat.py:158-164 re-runs initialize_pm): cheap, keeps plans. */
int hymd_ctx_set_box(hymd_ctx* ctx, const double box[3]);
/* New Hamiltonian parameters without rebuilding the context. */
int hymd_ctx_set_interaction(hymd_ctx* ctx, const double* A, const double* c, const double* m,
                             double sigma, double elec_conversion);

/* pm.decompose(positions[types==t]) for all t (main.py:977-980, 1007): bins the n local
 * particles by mesh-cell bin (two bins per 32 cells of a cell row).  d_pos is (n,3) row-major in the context dtype, d_types int32 (n),
 * d_charges (n) in the context dtype or NULL.  Positions may lie anywhere (they are wrapped
 * periodically).  With world_size > 1 the call is collective and does what pm.decompose +
 * Layout.exchange do in the reference: a particle whose cell belongs to another rank's slab (a bead
 * of a molecule that straddles a slab face, or a particle that drifted since the last
 * domain_decomposition) is stored into that rank's guest inbox over NVLink, painted and read out
 * there, and hymd_readout / hymd_pme_cycle bring its force back to the caller's row -- on the device,
 * without host synchronisation.  The guest buffers hold max(16384, n/4) particles per rank pair
 * (HYMD_B200_GUEST_CAPACITY overrides); exceeding them fails the next call with HYMD_ERR_CAPACITY.
 * Keeps no reference to the inputs after the stream work completes. */
int hymd_sort_particles(hymd_ctx* ctx, const void* d_pos, const int32_t* d_types,
                        const void* d_charges, int64_t n, void* stream);

/* Same, with flags.  HYMD_SORT_REUSE_ORDER: the per-index particle types are the same as in the
 * previous call and n is unchanged (true between consecutive MD steps: HyMD reads the types once
 * from the input file); the binning then starts from the previous bin order, which turns the
 * random scatter of a cold sort into nearly sequential traffic.  d_types may be NULL with this
 * flag.  If there is no usable previous order the flag is ignored (then d_types is required).
 * The result is identical either way. */
#define HYMD_SORT_REUSE_ORDER 1
int hymd_sort_particles_ex(hymd_ctx* ctx, const void* d_pos, const int32_t* d_types,
                           const void* d_charges, int64_t n, int flags, void* stream);
/* Forget the previous order (call when the caller's particle arrays were permuted or replaced). */
int hymd_ctx_reset_order(hymd_ctx* ctx);

/* Attach charges to an existing sort of the same positions (pm.decompose(positions) for the
 * charge layout, main.py:1007, without re-binning). */
int hymd_set_charges(hymd_ctx* ctx, const void* d_charges, void* stream);

/* pm.paint(...)/volume_per_cell for every type (field.py:574-575): deterministic CIC deposit
 * of the sorted particles into HYMD_FIELD_PHI[t]; includes the ghost-plane halo reduce. */
int hymd_paint(hymd_ctx* ctx, void* stream);

/* field.py:576-616: r2c, Gaussian filter, v_ext, second filter, -ik_d, c2r into
 * HYMD_FIELD_FORCE_MESH (+ ghost fill / halo fetch).  compute_potential != 0 also
 * materializes HYMD_FIELD_V_EXT (field.py:615-616) and the filtered HYMD_FIELD_PHI. */
int hymd_field_cycle(hymd_ctx* ctx, int compute_potential, void* stream);

/* update_field (hymd/field.py:428-616) in one call: hymd_sort_particles_ex + hymd_paint + hymd_field_cycle with the
 * same arguments and the same results.  On one GPU, for launch-bound systems, the step can be recorded once into a
 * CUDA graph and replayed with a single submission (hymd_ctx_set_graph): the recording is a stream capture of the
 * same host code, keyed on n, the flags, the current half of the record double buffer and a digest of the
 * configuration and of the context's buffers; d_pos may change from call to call (the recorded kernels read a
 * context-owned staging array that an ordinary copy on the caller's stream fills first).  Steps with
 * compute_potential != 0, with phase timing enabled, or on several slabs always take the ordinary launches. */
int hymd_update_cycle(hymd_ctx* ctx, const void* d_pos, const int32_t* d_types, const void* d_charges,
                      int64_t n, int flags, int compute_potential, void* stream);
/* Graph replay of hymd_update_cycle: 0 off, 1 on, -1 automatic (on when n <= 2^21 and the mesh has <= 2^21 cells).
 * The environment variable HYMD_B200_GRAPH (0 | 1 | auto) sets the initial mode of a context. */
int hymd_ctx_set_graph(hymd_ctx* ctx, int mode);
/* out = {steps replayed, graphs recorded, steps run eagerly by hymd_update_cycle, graphs alive}. */
int hymd_ctx_graph_stats(hymd_ctx* ctx, int64_t out[4]);

/* compute_field_force (field.py:152-200): d_force (n,3) row-major, caller particle order. */
int hymd_readout(hymd_ctx* ctx, void* d_force, void* stream);

/* update_field_force_q (field.py:241-403): uses the charges given to hymd_sort_particles;
 * writes d_elec_force (n,3).  want_psi != 0 also transforms psi back to real space. */
int hymd_pme_cycle(hymd_ctx* ctx, void* d_elec_force, int want_psi, void* stream);

/* Lazily materialize by-products the force path does not need every step:
 * filtered phi~ (field.py:578), phi_fourier (field.py:577), v_ext (field.py:616), and the
 * real-space electrostatic potential psi (field.py:377). */
int hymd_materialize(hymd_ctx* ctx, int want_phi, int want_phi_fourier, int want_v_ext,
                     int want_psi, void* stream);

/* compute_field_and_kinetic_energy (field.py:688-703) field terms: writes 2 doubles to the
 * HOST array out: {sum_cells w_0(phi~) dV, sum_cells 0.5 phi_q psi dV} for this slab (the
 * caller all-reduces and subtracts the self energy).  chi_upper is the T x T chi table,
 * kappa/rho0/a the Hamiltonian parameters, shift_a = 0 for SquaredPhi.  Synchronous. */
int hymd_field_energy(hymd_ctx* ctx, const double* chi, double kappa, double rho0, double a,
                      double out[2], void* stream);

/* comp_laplacian (field.py:406-425): HYMD_FIELD_PHI_LAPLACIAN[t][d] = c2r(-k_d^2 phi_fourier[t])
 * for the spectra of the last hymd_paint + hymd_field_cycle (cached until the next paint). */
int hymd_laplacian(hymd_ctx* ctx, void* stream);

/* Field terms of comp_pressure (pressure.py:105-127) for this slab.  With
 * V_t = c[t] + sum_j A[t*T+j] phi~_j (+ type_charges[t] psi when type_charges != NULL) writes 4
 * doubles to the HOST array out: {sum_cells sum_t V_t phi~_t, sum_cells sum_t V_t lap[t][d], d =
 * x,y,z}; the caller scales by dV/V (and sigma^2) and all-reduces.  Needs hymd_materialize
 * (filtered phi~), hymd_laplacian and, with type_charges, psi.  Synchronous. */
int hymd_field_pressure(hymd_ctx* ctx, const double* A, const double* c, const double* type_charges,
                        double out[4], void* stream);

/* Pointer + geometry of a context-owned buffer.  dims[3] = logical extents, pitch[3] = element
 * strides (in elements of the scalar or complex type). */
int hymd_get_field(hymd_ctx* ctx, int field_id, int t, int d, void** d_ptr, int64_t dims[3],
                   int64_t pitch[3]);

/* Synchronizes and reports {max particles per sort bin (>= per cell), local particles that are guests on another slab,
 * local particle count, distinct potential rows} of the last hymd_sort_particles. */
int hymd_ctx_status(hymd_ctx* ctx, int64_t out[4]);

/* Synchronizes the device and returns the sticky device-raised conditions of a multi-slab context
 * (guest capacity exceeded -> HYMD_ERR_CAPACITY, a peer-memory barrier timed out -> HYMD_ERR_NCCL);
 * the per-step entry points check the same flags without synchronizing. */
int hymd_ctx_check(hymd_ctx* ctx);

/* Which code paths this context runs (bench / test bookkeeping): out = {fused x-line kernel,
 * one-pass (y,z) plane transforms, slab pipeline, exchanges over NVLink peer memory}, 0 or 1. */
int hymd_ctx_paths(hymd_ctx* ctx, int32_t out[4]);

/* Number of CUDA kernels launched by this context since creation (bench bookkeeping). */
int64_t hymd_launch_count(hymd_ctx* ctx);

/* Per-phase device timing (bench bookkeeping; nothing in the reference corresponds).  While
 * enabled, every phase of the entry points above is bracketed by CUDA events on the caller's
 * stream.  hymd_ctx_get_timings synchronizes, ADDS the elapsed milliseconds and the number of
 * bracketed intervals of each phase since the last call into ms[HYMD_PHASE_COUNT] /
 * calls[HYMD_PHASE_COUNT] (the caller zeroes them), and recycles the events. */
typedef enum {
    HYMD_PHASE_SORT = 0,        /* cell binning: count + scan + scatter            */
    HYMD_PHASE_PAINT = 1,       /* CIC paint of all types                          */
    HYMD_PHASE_FFT_FWD = 2,     /* density r2c (cuFFT + pack/transposes)           */
    HYMD_PHASE_KSPACE = 3,      /* fused filter + potential + ik kernel            */
    HYMD_PHASE_FFT_INV = 4,     /* force-mesh c2r (cuFFT + pack/transposes)        */
    HYMD_PHASE_GHOST = 5,       /* periodic ghost planes / halo fetch of force meshes */
    HYMD_PHASE_READOUT = 6,     /* CIC gather of the forces                        */
    HYMD_PHASE_PME_PAINT = 7,
    HYMD_PHASE_PME_FFT = 8,
    HYMD_PHASE_PME_KSPACE = 9,
    HYMD_PHASE_PME_READOUT = 10,
    HYMD_PHASE_ALLTOALL = 11,   /* NCCL all-to-all of the slab FFT transposes (inside 2/4/8) */
    HYMD_PHASE_HALO = 12,       /* paint ghost-plane reduce (inside 1/7)           */
    HYMD_PHASE_MIGRATE = 13,    /* hymd_migrate                                    */
    HYMD_PHASE_BYPRODUCTS = 14, /* phi~ / v_ext / psi materialization              */
    HYMD_PHASE_COUNT = 16
} hymd_phase;
int hymd_ctx_set_timing(hymd_ctx* ctx, int enable);
int hymd_ctx_get_timings(hymd_ctx* ctx, double* ms, int64_t* calls);

/* domain_decomposition / layout.exchange (field.py:1115-1178, main.py:1171-1201): GPU-side
 * particle migration between slabs, in two phases so the caller can size its output arrays.
 *
 * hymd_migrate_plan: d_pos is the (n,3) array of ROUTING positions in the context dtype (the
 * particle positions, or for molecules the position of each molecule's first atom,
 * field.py:1156-1163).  Decides the destination slab of every row, exchanges the counts and
 * returns the new local row count in *n_new.  Synchronizes the stream.
 * hymd_migrate_apply: moves one per-particle array of row_bytes bytes per row: d_in has n rows,
 * d_out (distinct from d_in) receives n_new rows: first the rows that stay, in their original
 * relative order, then the arrivals from rank 0, 1, ...  Call once per array of the plan. */
int hymd_migrate_plan(hymd_ctx* ctx, const void* d_pos, int64_t n, int64_t* n_new, void* stream);
int hymd_migrate_apply(hymd_ctx* ctx, const void* d_in, void* d_out, int32_t row_bytes,
                       void* stream);

/* Layout.get_exchange_cost (pmesh; read by main.py:1304-1312 at verbose > 2): sent_to[q], q < world size, receives
 * the number of this rank's particles that the last hymd_sort_particles routed to slab q as guests (0 for q = rank
 * and on a single GPU).  Synchronizes the stream: a logging call, not part of the cycle. */
int hymd_exchange_cost(hymd_ctx* ctx, int64_t* sent_to, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row f3: general-Poisson-equation electrostatics, coulombtype "PIC_Spectral_GPE".
 * One GPU or several slabs (collective: every rank calls it, also one without particles).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t struct_size;                     /* = sizeof(hymd_gpe_params) */
    int32_t convergence_type;                /* config.convergence_type: 0 "max_diff" (default), 1 "csum",
                                                2 "euclidean_norm" (main.py:141-163) */
    int32_t max_iter;                        /* 100 in the reference (field.py:1037); <= 0 selects it */
    int32_t pad;
    double pol_mixing;                       /* config.pol_mixing (default 0.6) */
    double conv_crit;                        /* config.conv_crit (default 1e-6) */
    double coulomb_constant;                 /* config.coulomb_constant: eps0_inv = 4 pi k_e (field.py:1068) */
    double dielectric_type[HYMD_MAX_TYPES];  /* config.dielectric_type by type id (input_parser.py:1166-1177) */
    double type_charges[HYMD_MAX_TYPES];     /* config.type_charges */
} hymd_gpe_params;

/* update_field_force_q_GPE (hymd/field.py:964-1112) for the charges given to hymd_sort_particles and the
 * type densities of the last hymd_paint + hymd_field_cycle (they are materialized in filtered form
 * first).  Writes d_elec_force (n,3) in caller order (may be NULL) and the number of polarisation
 * iterations to *iterations (may be NULL).  Afterwards HYMD_FIELD_PHI_Q holds the filtered charge density
 * divided by the dielectric (field.py:1010, 1021), HYMD_FIELD_PSI the potential, HYMD_FIELD_GPE_* the
 * dielectric, |E|^2 and the per-type electrostatic potentials.  The context must have been created with
 * pme = 1.  Synchronizes the stream once per batch of four iterations (the convergence state lives on the device). */
int hymd_gpe_cycle(hymd_ctx* ctx, const hymd_gpe_params* params, void* d_elec_force, int32_t* iterations,
                   void* stream);
/* compute_field_energy_q_GPE (field.py:706-760): dV * eps_0 / 2 * sum_cells phi_eps |E|^2 of this slab to the
 * HOST double *out.  Synchronous. */
int hymd_gpe_energy(hymd_ctx* ctx, double coulomb_constant, double* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row f2: the caller side of the hot path kept on the device -- intramolecular
 * forces, velocity-Verlet / rRESPA updates and the CSVR thermostat -- so that positions and velocities
 * never leave HBM between field-force cycles.  These entry points are stateless with respect to
 * hymd_ctx (any stream, any number of contexts).
 * ---------------------------------------------------------------------------------------------- */

/* Device-resident term lists built from what prepare_bonds returns (hymd/force.py:573-728; HOST arrays
 * of local particle indices and parameters).  coeff4 is bonds_4_coeff (n4,6,5) row-major, type4 is
 * bonds_4_type: 0 cosine series, 1 combined bending-torsion (rows 4, 5 of its coefficients are the series of the
 * bending constant; hymd_bonded_set_last, hymd_bonded_dipoles below), 2 improper.  Synchronous.  Rebuild after
 * domain_decomposition permutes the particles (main.py:1239-1262 does the same with prepare_bonds).
 * A hymd_bonded owns one set of reduction scratch buffers: use it from one stream at a time. */
typedef struct hymd_bonded hymd_bonded;
int hymd_bonded_create(int64_t n_particles,
                       int64_t n2, const int32_t* a2, const int32_t* b2, const double* r0_2, const double* k_2,
                       int64_t n3, const int32_t* a3, const int32_t* b3, const int32_t* c3,
                       const double* t0_3, const double* k_3,
                       int64_t n4, const int32_t* a4, const int32_t* b4, const int32_t* c4, const int32_t* d4,
                       const double* coeff4, const int32_t* type4, hymd_bonded** out);
int hymd_bonded_destroy(hymd_bonded* b);

/* cbf / caf / cdf (hymd/compute_bond_forces.f90:1-61, compute_angle_forces.f90:1-93,
 * compute_dihedral_forces.f90:1-137; call sites main.py:841-887): kind = 2, 3 or 4 particles per term.
 * d_pos (n,3) and d_force (n,3) row-major in `dtype` (hymd_dtype); d_force is overwritten for every
 * particle (the Fortran zeroes f first).  d_out is a DEVICE array of 4 doubles that receives
 * {energy, pr_x, pr_y, pr_z} (bond_pr / angle_pr; zeros for dihedrals) -- no host synchronization. */
int hymd_bonded_forces(hymd_bonded* b, int kind, int dtype, const void* d_pos, const double box[3],
                       void* d_force, double* d_out, void* stream);
int64_t hymd_bonded_launch_count(hymd_bonded* b);

/* Dihedrals of dih_type 1 (compute_dihedral_forces.f90:77-112, dipole_reconstruction.f90:50-221).
 * hymd_bonded_set_last: last4 (HOST, n4) = bonds_4_last of prepare_bonds (force.py:569, the `bb_index` argument of
 *   cdf): 1 marks the last dihedral of a backbone, whose second angle b-c-d carries a bending term and a dipole as
 *   well.  All zeros until it is called.  Synchronous.
 * hymd_bonded_forces(kind = 4) includes the bending term (a second pass over the particles; topologies without
 *   dih_type 1 never launch it); hymd_bonded_inner_step runs such topologies through the per-particle variant of its
 *   kernel that carries the bending term, whatever hymd_bonded_set_cta selected.
 * hymd_bonded_dipoles: what cdf leaves in `dipoles` (n4,4,3) and `transfer_matrix` (n4,6,3,3) with dipole_flag = 1,
 *   row-major in `dtype` -- the reconstructed backbone dipole charges' positions (wrapped into the box) and the
 *   matrices that carry forces on them back to the beads; zeros for dihedrals of other types.
 * hymd_dipole_redistribute: dipole_forces_redistribution (hymd/force.py:855-880): d_f_dipoles (n4,4,3) are the
 *   electrostatic forces on the dipole charges (hymd_pme_cycle on the dipole positions, main.py:1060-1095);
 *   d_f_beads (n,3) is overwritten for every particle. */
int hymd_bonded_set_last(hymd_bonded* b, const int32_t* last4);
int hymd_bonded_dipoles(hymd_bonded* b, int dtype, const void* d_pos, const double box[3], void* d_dipoles,
                        void* d_transfer, void* stream);
int hymd_dipole_redistribute(hymd_bonded* b, int dtype, const void* d_f_dipoles, const void* d_transfer,
                             void* d_f_beads, void* stream);

/* Evaluation strategy of hymd_bonded_forces / hymd_bonded_inner_step.  0 (default): every particle
 * re-evaluates the terms it takes part in.  1: every CTA of 128 consecutive particles evaluates each
 * term touching it once into shared memory and the particles gather their slots (2-4x fewer
 * evaluations; the same forces).  2: as 1 with the bond / angle records stored inline in the CTA lists
 * and the CTA's own positions staged in shared memory.  3: inline records, positions from global memory
 * (not yet run on a GPU).  HYMD_ERR_CAPACITY if a CTA's term list does not fit in shared
 * memory.  The environment variable HYMD_B200_BONDED_CTA=1 selects it at creation. */
int hymd_bonded_set_cta(hymd_bonded* b, int enable);

/* Arithmetic of the bond and angle terms in hymd_bonded_inner_step for float32 positions with per-particle
 * evaluation.  0 (default): double, term by term like the Fortran.  1: single precision with formulas that
 * stay accurate there (atan2 of cross and dot products instead of acos, no cancelling subtractions):
 * <= 1e-5 of the largest force against the double path (measured 3e-7 ... 1.3e-6 on the CPU), i.e. inside the
 * fp32 build's tolerance, for a fraction of the fp64 instruction count.  HYMD_B200_BONDED_F32MATH=1 selects
 * it at creation.  Never run on a GPU yet. */
int hymd_bonded_set_math(hymd_bonded* b, int f32math);

/* One fused inner rRESPA step (main.py:829-893) in a single pass over the particles:
 *   F = bond + angle + dihedral forces at d_pos_in (every kind rounded to `dtype` like the f arrays),
 *   n_kicks (0, 1 or 2) times  v += 0.5*kick_dt * F / mass   -- the closing kick of the previous inner
 *   step and the opening kick of the next one use the same forces and stay two separate roundings --
 *   then, if d_pos_out != NULL,  d_pos_out = mod(d_pos_in + drift_dt * v, box)  (a different buffer:
 *   neighbours still read d_pos_in).
 * An inner loop of r steps is r+1 launches: (1 kick, drift), r-1 x (2 kicks, drift), (1 kick, no drift).
 * d_force_out: NULL or HOST array of 3 device pointers (bond, angle, dihedral forces; entries may be
 * NULL) for callers that want the per-kind arrays; d_out: NULL or DEVICE array of 12 doubles =
 * {energy, pr_x, pr_y, pr_z} x {bonds, angles, dihedrals} at d_pos_in. */
int hymd_bonded_inner_step(hymd_bonded* b, int dtype, const void* d_pos_in, void* d_pos_out, void* d_vel,
                           const double box[3], double mass, double kick_dt, int n_kicks, double drift_dt,
                           void* const* d_force_out, double* d_out, void* stream);

/* integrate_velocity / integrate_position (hymd/integrator.py:9-75) fused over the arrays:
 *   sequential == 0:  v += 0.5*kick_dt * (f_0 + ... + f_{n_forces-1}) / mass      (main.py:830-834, 889-893)
 *   sequential == 1:  v += 0.5*kick_dt * f_k / mass  for k = 0, 1, ... in turn    (main.py:803-827, 1144-1169)
 * and, if d_pos != NULL,  x = mod(x + drift_dt * v, box)                           (main.py:835-837)
 * in the same pass.  d_forces is a HOST array of n_forces (<= 8) device pointers; n_forces may be 0
 * (drift only). */
int hymd_md_kick_drift(int dtype, void* d_vel, void* d_pos, const void* const* d_forces, int n_forces,
                       int sequential, double mass, double kick_dt, double drift_dt, const double box[3],
                       int64_t n, void* stream);

/* Velocity moments {count, sum vx, sum vy, sum vz, sum |v|^2} of the particles with d_group[i] == group
 * (every particle if d_group == NULL or group < 0) -> d_out[0..4], and of all n particles ->
 * d_out[5..9] (DEVICE doubles; the caller all-reduces them across ranks where the reference calls
 * comm.allreduce: thermostat.py:13, 184-191; kinetic energy field.py:695).  d_scratch holds
 * hymd_velocity_moments_scratch_doubles() doubles.  Fixed summation order. */
int hymd_velocity_moments(int dtype, const void* d_vel, const int32_t* d_group, int group, int64_t n,
                          double* d_scratch, double* d_out, void* stream);
int64_t hymd_velocity_moments_scratch_doubles(void);

/* csvr_thermostat for ONE coupling group (hymd/thermostat.py:184-219) given the reduced moments:
 * kT15 = 1.5 * gas_constant * target_temperature, c = exp(-time_step*respa_inner/tau), R and SNf the
 * Gaussian and chi-squared draws of this group (drawn on the host from the caller's prng in the
 * reference's order).  remove_com != 0 and more than one particle in the group: the group is rescaled
 * about its centre-of-mass velocity; otherwise, like the reference, the kinetic energy of ALL particles
 * is used and ALL velocities are rescaled.  Adds the thermostat work dK to d_work[0] (device double;
 * may be NULL). */
int hymd_csvr_apply(int dtype, void* d_vel, const int32_t* d_group, int group, int64_t n,
                    const double* d_moments, double mass, double kT15, double c, double R, double SNf,
                    int remove_com, double* d_work, void* stream);

/* cancel_com_momentum (hymd/thermostat.py:12-15): v -= (sum v over all particles) / n_particles with
 * the sums taken from d_moments[6..8]. */
int hymd_cancel_com(int dtype, void* d_vel, int64_t n, const double* d_moments, double n_particles,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HYMD_B200_H */

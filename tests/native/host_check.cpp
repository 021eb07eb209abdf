// TEST INFRASTRUCTURE: runs the __host__ __device__ per-particle functions of
// hymd_b200/csrc/bonded.cuh and md.cuh -- the same source the sm_100a kernels compile -- on the CPU,
// so tests/test_native_host_check.py can compare the device arithmetic (term lists, slot logic,
// minimum image, cosine series, CSVR scale) with the oracle in the GPU-less container.  Built by the
// test with g++; never linked into libhymd_b200.so and never used by the product.
#include <stdlib.h>

#include "../../hymd_b200/csrc/bonded.cuh"
#include "../../hymd_b200/csrc/md.cuh"

using namespace hymd;

template <typename real, int KIND>
static void run_kind(const real* pos, Vec3d box, long long n, const std::vector<uint32_t>& start,
                     const std::vector<uint32_t>& refs, const std::vector<int32_t>& idx, const double* par,
                     const int32_t* dtype, real* force, double* out4) {
    out4[0] = out4[1] = out4[2] = out4[3] = 0.0;
    for (long long p = 0; p < n; ++p) {
        BondAcc a = particle_terms<real, KIND>(p, pos, box, start.data(), refs.data(), idx.data(), par, dtype);
        force[3 * p] = (real)a.f.x; force[3 * p + 1] = (real)a.f.y; force[3 * p + 2] = (real)a.f.z;
        out4[0] += a.e; out4[1] += a.pr.x; out4[2] += a.pr.y; out4[3] += a.pr.z;
    }
}

extern "C" int host_bonded(int kind, int f64, const void* pos, const double* box, long long n, long long nt,
                           const int32_t* a, const int32_t* b, const int32_t* c, const int32_t* d,
                           const double* par, const int32_t* dtype, void* force, double* out4) {
    const int32_t* index[4] = {a, b, c, d};
    std::vector<uint32_t> start, refs;
    if (!build_particle_csr(n, nt, kind, index, start, refs)) return -1;
    std::vector<int32_t> idx((size_t)nt * 4, 0);
    for (long long t = 0; t < nt; ++t)
        for (int s = 0; s < kind; ++s) idx[4 * t + s] = index[s][t];
    const Vec3d bx = {box[0], box[1], box[2]};
#define RUN(real, K) run_kind<real, K>((const real*)pos, bx, n, start, refs, idx, par, dtype, (real*)force, out4)
    if (f64) { if (kind == 2) RUN(double, 2); else if (kind == 3) RUN(double, 3); else RUN(double, 4); }
    else     { if (kind == 2) RUN(float, 2);  else if (kind == 3) RUN(float, 3);  else RUN(float, 4); }
#undef RUN
    return 0;
}

// dtype-1 dihedrals: what the device adds on top of host_bonded(kind = 4) -- the bending term's forces and energy
// (cbt_eval, gathered per particle like the device's cbt pass), the dipoles and transfer matrices (dipole_term) and
// the redistribution of dipole forces to the beads (redistribute_term, gathered per particle)
template <typename real>
static void run_cbt(const real* pos, Vec3d bx, long long n, long long nt, const std::vector<uint32_t>& start,
                    const std::vector<uint32_t>& refs, const std::vector<int32_t>& idx, const double* par,
                    const int32_t* dtype, const int32_t* last, real* force_add, double* energy, real* dipoles,
                    real* transfer, const real* f_dipoles, real* f_beads) {
    *energy = 0.0;
    for (long long p = 0; p < n; ++p) {
        Vec3d acc = {0, 0, 0}, accb = {0, 0, 0};
        for (uint32_t k = start[p]; k < start[p + 1]; ++k) {
            const long long t = refs[k] >> 2;
            const int slot = refs[k] & 3;
            if (dtype[t] != 1) continue;
            const int32_t* ix = &idx[4 * t];
            Vec3d out[4];
            double e;
            cbt_eval(pos, bx, ix[0], ix[1], ix[2], ix[3], par + (long long)DIH_ROWS * DIH_COLS * t, last[t], out, e);
            acc = acc + out[slot];
            if (slot == 0) *energy += e;
            if (f_dipoles) accb = accb + redistribute_term<real>(slot, last[t], f_dipoles + 12 * t, transfer + 54 * t);
        }
        force_add[3 * p] = (real)acc.x; force_add[3 * p + 1] = (real)acc.y; force_add[3 * p + 2] = (real)acc.z;
        if (f_beads) { f_beads[3 * p] = (real)accb.x; f_beads[3 * p + 1] = (real)accb.y; f_beads[3 * p + 2] = (real)accb.z; }
    }
}

extern "C" int host_cbt(int f64, const void* pos, const double* box, long long n, long long nt, const int32_t* a,
                        const int32_t* b, const int32_t* c, const int32_t* d, const double* par, const int32_t* dtype,
                        const int32_t* last, void* force_add, double* energy, void* dipoles, void* transfer,
                        const void* f_dipoles, void* f_beads) {
    const int32_t* index[4] = {a, b, c, d};
    std::vector<uint32_t> start, refs;
    if (!build_particle_csr(n, nt, 4, index, start, refs)) return -1;
    std::vector<int32_t> idx((size_t)nt * 4, 0);
    for (long long t = 0; t < nt; ++t)
        for (int s = 0; s < 4; ++s) idx[4 * t + s] = index[s][t];
    const Vec3d bx = {box[0], box[1], box[2]};
    // dipoles first: the redistribution reads the transfer matrices
    for (long long t = 0; t < nt; ++t) {
        const int32_t* ix = &idx[4 * t];
        if (f64) dipole_term<double>((const double*)pos, bx, ix[0], ix[1], ix[2], ix[3], par + (long long)DIH_ROWS * DIH_COLS * t,
                                     dtype[t], last[t], (double*)dipoles + 12 * t, (double*)transfer + 54 * t);
        else dipole_term<float>((const float*)pos, bx, ix[0], ix[1], ix[2], ix[3], par + (long long)DIH_ROWS * DIH_COLS * t,
                                dtype[t], last[t], (float*)dipoles + 12 * t, (float*)transfer + 54 * t);
    }
    if (f64) run_cbt<double>((const double*)pos, bx, n, nt, start, refs, idx, par, dtype, last, (double*)force_add, energy,
                             (double*)dipoles, (double*)transfer, (const double*)f_dipoles, (double*)f_beads);
    else run_cbt<float>((const float*)pos, bx, n, nt, start, refs, idx, par, dtype, last, (float*)force_add, energy,
                        (float*)dipoles, (float*)transfer, (const float*)f_dipoles, (float*)f_beads);
    return 0;
}

template <typename real>
static void kd(real* vel, real* pos, const void* const* forces, int nf, int sequential, double mass,
               double kick_dt, double drift_dt, const double* box, long long n) {
    for (long long i = 0; i < 3 * n; ++i) {
        real v = vel[i];
        if (nf > 0) {
            real ft[MD_MAX_FORCES];
            for (int k = 0; k < nf; ++k) ft[k] = ((const real*)forces[k])[i];
            if (sequential) for (int k = 0; k < nf; ++k) v = kick(v, &ft[k], 1, (real)mass, (real)(0.5 * kick_dt));
            else v = kick(v, ft, nf, (real)mass, (real)(0.5 * kick_dt));
            vel[i] = v;
        }
        if (pos) pos[i] = drift_wrap(pos[i], v, (real)drift_dt, (real)box[i % 3]);
    }
}

extern "C" int host_kick_drift(int f64, void* vel, void* pos, const void* const* forces, int nf,
                               int sequential, double mass, double kick_dt, double drift_dt,
                               const double* box, long long n) {
    if (f64) kd<double>((double*)vel, (double*)pos, forces, nf, sequential, mass, kick_dt, drift_dt, box, n);
    else kd<float>((float*)vel, (float*)pos, forces, nf, sequential, mass, kick_dt, drift_dt, box, n);
    return 0;
}

template <typename real>
static void csvr(real* vel, const int32_t* group, int g, long long n, double mass, double kT15, double c,
                 double R, double SNf, int remove_com, double* work) {
    double mom[2 * MOM] = {0};
    for (long long i = 0; i < n; ++i) {
        const double vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
        const bool in_g = (!group || g < 0) ? true : group[i] == g;
        if (in_g) { mom[0] += 1; mom[1] += vx; mom[2] += vy; mom[3] += vz; mom[4] += vx * vx + vy * vy + vz * vz; }
        mom[5] += 1; mom[6] += vx; mom[7] += vy; mom[8] += vz; mom[9] += vx * vx + vy * vy + vz * vz;
    }
    double dK;
    const CsvrScale s = csvr_scale(mom, mass, kT15, c, R, SNf, remove_com, &dK);
    for (long long i = 0; i < n; ++i)
        csvr_apply_particle(vel + 3 * i, s, (!group || g < 0) ? true : group[i] == g);
    work[0] += dK;
}

extern "C" int host_csvr(int f64, void* vel, const int32_t* group, int g, long long n, double mass,
                         double kT15, double c, double R, double SNf, int remove_com, double* work) {
    if (f64) csvr<double>((double*)vel, group, g, n, mass, kT15, c, R, SNf, remove_com, work);
    else csvr<float>((float*)vel, group, g, n, mass, kT15, c, R, SNf, remove_com, work);
    return 0;
}

// ---- the row-f2 C ABI (include/hymd_b200.h) emulated on HOST memory ------------------------------
// Same entry-point names and signatures as libhymd_b200.so, so that tests/test_md_host_emulation.py
// can run the Python host layer (hymd_b200/force.py, thermostat.py, md.py) and the GPU test bodies
// on CPU tensors in the GPU-less container.  The loops call the same per-particle functions the
// kernels call; only launch geometry and the parallel reductions differ.
struct HostBonded {
    long long n;
    std::vector<uint32_t> start[3], refs[3];
    std::vector<int32_t> idx[3], dtype, last;
    std::vector<double> par[3];
    std::vector<uint32_t> cta_start[3], cta_terms[3], lrefs[3];
    std::vector<TermRec> rec[2];
    int max_terms[3];
    int use_cta;
    int f32math;
    int tile;                   // particles per CTA (HYMD_B200_BONDED_TILE, default BONDED_THREADS = 128)
    long long launches;
};
constexpr int HOST_THREADS = 128;   // BONDED_THREADS of csrc/bonded.cu

static bool host_upload(HostBonded* b, int kind, long long nt, int slots, const int32_t* const* index,
                        const double* par, size_t per_term) {
    if (!build_particle_csr(b->n, nt, slots, index, b->start[kind], b->refs[kind])) return false;
    b->idx[kind].assign((size_t)nt * 4, 0);
    for (long long t = 0; t < nt; ++t)
        for (int s = 0; s < slots; ++s) b->idx[kind][4 * t + s] = index[s][t];
    b->par[kind].assign(par, par + (size_t)nt * per_term);
    build_cta_lists(b->n, nt, slots, index, b->tile, b->start[kind], b->cta_start[kind], b->cta_terms[kind],
                    b->lrefs[kind], b->max_terms[kind]);
    if (kind < 2) build_cta_records(b->cta_terms[kind], b->idx[kind].data(), b->par[kind].data(), b->rec[kind]);
    return true;
}

extern "C" int hymd_bonded_create(int64_t n_particles, int64_t n2, const int32_t* a2, const int32_t* b2,
                                  const double* r0_2, const double* k_2, int64_t n3, const int32_t* a3,
                                  const int32_t* b3, const int32_t* c3, const double* t0_3, const double* k_3,
                                  int64_t n4, const int32_t* a4, const int32_t* b4, const int32_t* c4,
                                  const int32_t* d4, const double* coeff4, const int32_t* type4, void** out) {
    for (int64_t t = 0; t < n4; ++t)
        if (type4[t] < 0 || type4[t] > 2) return -1;
    HostBonded* b = new HostBonded();
    b->n = n_particles;
    b->launches = 0;
    b->use_cta = 0;
    b->f32math = 0;
    b->tile = HOST_THREADS;
    if (const char* env = getenv("HYMD_B200_BONDED_TILE")) {
        b->tile = atoi(env);
        if (b->tile < HOST_THREADS || b->tile > 2048 || b->tile % HOST_THREADS != 0) { delete b; return -1; }
    }
    std::vector<double> p2((size_t)n2 * 2), p3((size_t)n3 * 2);
    for (int64_t t = 0; t < n2; ++t) { p2[2 * t] = r0_2[t]; p2[2 * t + 1] = k_2[t]; }
    for (int64_t t = 0; t < n3; ++t) { p3[2 * t] = t0_3[t]; p3[2 * t + 1] = k_3[t]; }
    const int32_t* i2[2] = {a2, b2};
    const int32_t* i3[3] = {a3, b3, c3};
    const int32_t* i4[4] = {a4, b4, c4, d4};
    if (!host_upload(b, 0, n2, 2, i2, p2.data(), 2) || !host_upload(b, 1, n3, 3, i3, p3.data(), 2) ||
        !host_upload(b, 2, n4, 4, i4, coeff4, (size_t)DIH_ROWS * DIH_COLS)) {
        delete b;
        return -1;
    }
    b->dtype.assign(type4, type4 + n4);
    b->last.assign((size_t)n4, 0);
    *out = b;
    return 0;
}

extern "C" int hymd_bonded_set_last(void* h, const int32_t* last4) {
    HostBonded* b = (HostBonded*)h;
    b->last.assign(last4, last4 + b->dtype.size());
    return 0;
}

// the bending pass of csrc/bonded.cu (cbt_kernel + cbt_final_kernel) on top of a dihedral force array
template <typename real>
static void host_cbt_pass(HostBonded* b, const real* pos, Vec3d bx, real* force, double* out) {
    bool any = false;
    for (int32_t t : b->dtype) any |= t == 1;
    if (!any) return;
    for (long long p = 0; p < b->n; ++p) {
        Vec3d acc = {0, 0, 0};
        bool hit = false;
        for (uint32_t k = b->start[2][p]; k < b->start[2][p + 1]; ++k) {
            const long long t = b->refs[2][k] >> 2;
            const int slot = b->refs[2][k] & 3;
            if (b->dtype[t] != 1) continue;
            const int32_t* ix = &b->idx[2][4 * t];
            Vec3d o[4];
            double e;
            cbt_eval(pos, bx, ix[0], ix[1], ix[2], ix[3], b->par[2].data() + (long long)DIH_ROWS * DIH_COLS * t, b->last[t], o, e);
            acc = acc + o[slot];
            if (slot == 0) out[0] += e;
            hit = true;
        }
        if (hit) {
            force[3 * p] = (real)((double)force[3 * p] + acc.x);
            force[3 * p + 1] = (real)((double)force[3 * p + 1] + acc.y);
            force[3 * p + 2] = (real)((double)force[3 * p + 2] + acc.z);
        }
    }
}

extern "C" int hymd_bonded_dipoles(void* h, int dtype, const void* pos, const double* box, void* dipoles, void* transfer,
                                   void* stream) {
    HostBonded* b = (HostBonded*)h;
    const Vec3d bx = {box[0], box[1], box[2]};
    for (size_t t = 0; t < b->dtype.size(); ++t) {
        const int32_t* ix = &b->idx[2][4 * t];
        const double* par = b->par[2].data() + (long long)DIH_ROWS * DIH_COLS * t;
        if (dtype == 1) dipole_term<double>((const double*)pos, bx, ix[0], ix[1], ix[2], ix[3], par, b->dtype[t], b->last[t],
                                            (double*)dipoles + 12 * t, (double*)transfer + 54 * t);
        else dipole_term<float>((const float*)pos, bx, ix[0], ix[1], ix[2], ix[3], par, b->dtype[t], b->last[t],
                                (float*)dipoles + 12 * t, (float*)transfer + 54 * t);
    }
    return 0;
}

extern "C" int hymd_dipole_redistribute(void* h, int dtype, const void* f_dipoles, const void* transfer, void* f_beads,
                                        void* stream) {
    HostBonded* b = (HostBonded*)h;
    for (long long p = 0; p < b->n; ++p) {
        Vec3d acc = {0, 0, 0};
        for (uint32_t k = b->start[2][p]; k < b->start[2][p + 1]; ++k) {
            const long long t = b->refs[2][k] >> 2;
            if (b->dtype[t] != 1) continue;
            const int slot = b->refs[2][k] & 3;
            if (dtype == 1) acc = acc + redistribute_term<double>(slot, b->last[t], (const double*)f_dipoles + 12 * t, (const double*)transfer + 54 * t);
            else acc = acc + redistribute_term<float>(slot, b->last[t], (const float*)f_dipoles + 12 * t, (const float*)transfer + 54 * t);
        }
        if (dtype == 1) { ((double*)f_beads)[3 * p] = acc.x; ((double*)f_beads)[3 * p + 1] = acc.y; ((double*)f_beads)[3 * p + 2] = acc.z; }
        else { ((float*)f_beads)[3 * p] = (float)acc.x; ((float*)f_beads)[3 * p + 1] = (float)acc.y; ((float*)f_beads)[3 * p + 2] = (float)acc.z; }
    }
    return 0;
}

extern "C" int hymd_bonded_destroy(void* b) { delete (HostBonded*)b; return 0; }
extern "C" int64_t hymd_bonded_launch_count(void* b) { return ((HostBonded*)b)->launches; }

template <typename real>
static void inner(HostBonded* b, int kind_mask, const real* x_in, real* x_out, real* vel, Vec3d box, double mass,
                  double kick_dt, int n_kicks, double drift_dt, void* const* f_out, double* out12);
extern "C" int host_inner_step_f32(void* h, const float* x_in, float* x_out, float* vel, const double* box,
                                   double mass, double kick_dt, int n_kicks, double drift_dt, void* const* f_out,
                                   double* out12);

extern "C" int hymd_bonded_forces(void* h, int kind, int dtype, const void* pos, const double* box,
                                  void* force, double* out, void* stream) {
    HostBonded* b = (HostBonded*)h;
    const int k = kind - 2;
    const Vec3d bx = {box[0], box[1], box[2]};
    b->launches += 2;
    if (b->f32math && dtype == 0 && !b->use_cta && k < 2) {
        void* fo1[3] = {nullptr, nullptr, nullptr};
        fo1[k] = force;
        double o12[12];
        host_inner_step_f32(h, (const float*)pos, nullptr, nullptr, box, 1.0, 0.0, 0, 0.0, fo1, o12);
        // (the fused f32 evaluator with the other kinds present but unrequested gives the same kind-k result)
        for (int j = 0; j < 4; ++j) out[j] = o12[4 * k + j];
        return 0;
    }
    if (b->use_cta) {
        void* fo[3] = {nullptr, nullptr, nullptr};
        fo[k] = force;
        double out12[12];
        if (dtype == 1) inner<double>(b, 1 << k, (const double*)pos, nullptr, nullptr, bx, 1.0, 0.0, 0, 0.0, fo, out12);
        else inner<float>(b, 1 << k, (const float*)pos, nullptr, nullptr, bx, 1.0, 0.0, 0, 0.0, fo, out12);
        for (int j = 0; j < 4; ++j) out[j] = out12[4 * k + j];
        if (kind == 4) { if (dtype == 1) host_cbt_pass<double>(b, (const double*)pos, bx, (double*)force, out); else host_cbt_pass<float>(b, (const float*)pos, bx, (float*)force, out); }
        return 0;
    }
#define RUN(real, K) run_kind<real, K>((const real*)pos, bx, b->n, b->start[k], b->refs[k], b->idx[k], \
                                       b->par[k].data(), b->dtype.data(), (real*)force, out)
    if (dtype == 1) { if (kind == 2) RUN(double, 2); else if (kind == 3) RUN(double, 3); else RUN(double, 4); }
    else            { if (kind == 2) RUN(float, 2);  else if (kind == 3) RUN(float, 3);  else RUN(float, 4); }
#undef RUN
    if (kind == 4) { if (dtype == 1) host_cbt_pass<double>(b, (const double*)pos, bx, (double*)force, out); else host_cbt_pass<float>(b, (const float*)pos, bx, (float*)force, out); }
    return 0;
}

extern "C" int hymd_md_kick_drift(int dtype, void* vel, void* pos, const void* const* forces, int nf,
                                  int sequential, double mass, double kick_dt, double drift_dt,
                                  const double* box, int64_t n, void* stream) {
    if (nf < 0 || nf > MD_MAX_FORCES) return -1;
    return host_kick_drift(dtype == 1, vel, pos, forces, nf, sequential, mass, kick_dt, drift_dt, box, n);
}

extern "C" int64_t hymd_velocity_moments_scratch_doubles(void) { return 16; }

template <typename real>
static void moments(const real* vel, const int32_t* group, int g, long long n, double* mom) {
    for (int k = 0; k < 2 * MOM; ++k) mom[k] = 0.0;
    for (long long i = 0; i < n; ++i) {
        const double vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
        const bool in_g = (!group || g < 0) ? true : group[i] == g;
        if (in_g) { mom[0] += 1; mom[1] += vx; mom[2] += vy; mom[3] += vz; mom[4] += vx * vx + vy * vy + vz * vz; }
        mom[5] += 1; mom[6] += vx; mom[7] += vy; mom[8] += vz; mom[9] += vx * vx + vy * vy + vz * vz;
    }
}

extern "C" int hymd_velocity_moments(int dtype, const void* vel, const int32_t* group, int g, int64_t n,
                                     double* scratch, double* out, void* stream) {
    if (dtype == 1) moments<double>((const double*)vel, group, g, n, out);
    else moments<float>((const float*)vel, group, g, n, out);
    return 0;
}

template <typename real>
static void apply(real* vel, const int32_t* group, int g, long long n, const double* mom, double mass,
                  double kT15, double c, double R, double SNf, int remove_com, double* work) {
    double dK;
    const CsvrScale s = csvr_scale(mom, mass, kT15, c, R, SNf, remove_com, &dK);
    for (long long i = 0; i < n; ++i)
        csvr_apply_particle(vel + 3 * i, s, (!group || g < 0) ? true : group[i] == g);
    if (work) work[0] += dK;
}

extern "C" int hymd_csvr_apply(int dtype, void* vel, const int32_t* group, int g, int64_t n, const double* mom,
                               double mass, double kT15, double c, double R, double SNf, int remove_com,
                               double* work, void* stream) {
    if (dtype == 1) apply<double>((double*)vel, group, g, n, mom, mass, kT15, c, R, SNf, remove_com, work);
    else apply<float>((float*)vel, group, g, n, mom, mass, kT15, c, R, SNf, remove_com, work);
    return 0;
}

extern "C" int hymd_cancel_com(int dtype, void* vel, int64_t n, const double* mom, double n_particles,
                               void* stream) {
    const double cx = mom[MOM + 1] / n_particles, cy = mom[MOM + 2] / n_particles, cz = mom[MOM + 3] / n_particles;
    for (long long i = 0; i < n; ++i) {
        if (dtype == 1) {
            double* v = (double*)vel + 3 * i;
            v[0] -= cx; v[1] -= cy; v[2] -= cz;
        } else {
            float* v = (float*)vel + 3 * i;
            v[0] = (float)((double)v[0] - cx); v[1] = (float)((double)v[1] - cy); v[2] = (float)((double)v[2] - cz);
        }
    }
    return 0;
}

template <typename real>
static void inner(HostBonded* b, int kind_mask, const real* x_in, real* x_out, real* vel, Vec3d box, double mass,
                  double kick_dt, int n_kicks, double drift_dt, void* const* f_out, double* out12) {
    TermLists t;
    CtaLists c;
    for (int k = 0; k < 3; ++k) {
        t.start[k] = b->start[k].data(); t.refs[k] = b->refs[k].data(); t.idx[k] = b->idx[k].data();
        t.par[k] = b->par[k].data();
        t.n_terms[k] = (kind_mask >> k) & 1 ? (long long)(b->idx[k].size() / 4) : 0;
        c.cta_start[k] = b->cta_start[k].data(); c.cta_terms[k] = b->cta_terms[k].data();
        c.lrefs[k] = b->lrefs[k].data(); c.max_terms[k] = b->max_terms[k];
    }
    t.dih_type = b->dtype.data();
    real* fo[3] = {f_out ? (real*)f_out[0] : nullptr, f_out ? (real*)f_out[1] : nullptr,
                   f_out ? (real*)f_out[2] : nullptr};
    double acc12[12] = {0};
    if (b->use_cta) {
        // the kernel's structure: per CTA, phase 1 (threads stride over the CTA's terms), barrier, phase 2
        std::vector<double> sm((size_t)3 * c.max_terms[0] + 6 * c.max_terms[1] + 12 * c.max_terms[2] + 1);
        const int HOST_CTA = HOST_THREADS;      // threads striding over the CTA's terms
        const long long n_cta = (b->n + b->tile - 1) / b->tile;
        for (long long cta = 0; cta < n_cta; ++cta) {
            const long long p0 = cta * b->tile, p1 = p0 + b->tile < b->n ? p0 + b->tile : b->n;
            for (size_t i = 0; i < sm.size(); ++i) sm[i] = -777.0;      // stale data must never be read
            if (b->use_cta == 3) {      // mode 3: inline records, positions from global memory
                CtaRecs rc;
                rc.rec[0] = b->rec[0].data();
                rc.rec[1] = b->rec[1].data();
                for (int tid = 0; tid < HOST_CTA; ++tid)
                    cta2_eval_terms<real, const real*>(tid, HOST_CTA, cta, p0, p1, x_in, box, t, c, rc, sm.data(), acc12);
            } else if (b->use_cta == 2) {      // mode 2: own positions staged, inline records
                std::vector<real> tile((size_t)3 * b->tile, (real)-555);
                for (long long i = 0; i < 3 * (p1 - p0); ++i) tile[i] = x_in[3 * p0 + i];
                const PosTile<real> xt = {x_in, tile.data(), p0, p1};
                CtaRecs rc;
                rc.rec[0] = b->rec[0].data();
                rc.rec[1] = b->rec[1].data();
                for (int tid = 0; tid < HOST_CTA; ++tid)
                    cta2_eval_terms<real, PosTile<real>>(tid, HOST_CTA, cta, p0, p1, xt, box, t, c, rc, sm.data(), acc12);
            } else
                for (int tid = 0; tid < HOST_CTA; ++tid)
                    cta_eval_terms<real>(tid, HOST_CTA, cta, p0, p1, x_in, box, t, c, sm.data(), acc12);
            for (long long p = p0; p < p1; ++p) {
                BondAcc acc[3];
                cta_gather_particle(p, t, c, sm.data(), acc);
                finish_particle<real>(p, x_in, x_out, vel, box, (real)mass, (real)(0.5 * kick_dt), n_kicks,
                                      (real)drift_dt, fo, acc);
            }
        }
    } else {
        for (long long p = 0; p < b->n; ++p) {
            BondAcc acc[3];
            inner_step_particle<real>(p, x_in, x_out, vel, box, t, (real)mass, (real)(0.5 * kick_dt), n_kicks,
                                      (real)drift_dt, fo, acc);
            for (int k = 0; k < 3; ++k) {
                acc12[4 * k] += acc[k].e; acc12[4 * k + 1] += acc[k].pr.x; acc12[4 * k + 2] += acc[k].pr.y;
                acc12[4 * k + 3] += acc[k].pr.z;
            }
        }
    }
    if (out12) for (int k = 0; k < 12; ++k) out12[k] = acc12[k];
}

template <typename real>
static void inner_cbt(HostBonded* b, const real* x_in, real* x_out, real* vel, const double* box, double mass,
                      double kick_dt, int n_kicks, double drift_dt, void* const* f_out, double* out12) {
    TermLists t;
    for (int k = 0; k < 3; ++k) {
        t.start[k] = b->start[k].data(); t.refs[k] = b->refs[k].data(); t.idx[k] = b->idx[k].data();
        t.par[k] = b->par[k].data(); t.n_terms[k] = (long long)(b->idx[k].size() / 4);
    }
    t.dih_type = b->dtype.data();
    t.dih_last = b->last.data();
    const Vec3d bx = {box[0], box[1], box[2]};
    real* fo[3] = {f_out ? (real*)f_out[0] : nullptr, f_out ? (real*)f_out[1] : nullptr, f_out ? (real*)f_out[2] : nullptr};
    double acc12[12] = {0};
    for (long long p = 0; p < b->n; ++p) {
        BondAcc acc[3];
        inner_step_particle_cbt<real>(p, x_in, x_out, vel, bx, t, (real)mass, (real)(0.5 * kick_dt), n_kicks,
                                      (real)drift_dt, fo, acc);
        for (int k = 0; k < 3; ++k) {
            acc12[4 * k] += acc[k].e; acc12[4 * k + 1] += acc[k].pr.x; acc12[4 * k + 2] += acc[k].pr.y;
            acc12[4 * k + 3] += acc[k].pr.z;
        }
    }
    if (out12) for (int k = 0; k < 12; ++k) out12[k] = acc12[k];
}

extern "C" int hymd_bonded_set_cta(void* h, int enable) { ((HostBonded*)h)->use_cta = (enable == 2 || enable == 3) ? enable : (enable ? 1 : 0); return 0; }

extern "C" int host_inner_step_f32(void* h, const float* x_in, float* x_out, float* vel, const double* box,
                                   double mass, double kick_dt, int n_kicks, double drift_dt, void* const* f_out,
                                   double* out12);
extern "C" int hymd_bonded_set_math(void* h, int f32math) { ((HostBonded*)h)->f32math = f32math ? 1 : 0; return 0; }

extern "C" int hymd_bonded_inner_step(void* h, int dtype, const void* x_in, void* x_out, void* vel,
                                      const double* box, double mass, double kick_dt, int n_kicks,
                                      double drift_dt, void* const* f_out, double* out12, void* stream) {
    HostBonded* b = (HostBonded*)h;
    if (n_kicks < 0 || n_kicks > 2 || x_in == x_out) return -1;
    bool cbt = false;
    for (int32_t t : b->dtype) cbt = cbt || t == 1;
    if (cbt) {      // dtype-1 dihedrals: the per-particle variant that carries the bending term (inner_step_kernel<real, true>)
        b->launches += out12 ? 2 : 1;
        if (dtype == 1) inner_cbt<double>(b, (const double*)x_in, (double*)x_out, (double*)vel, box, mass, kick_dt, n_kicks, drift_dt, f_out, out12);
        else inner_cbt<float>(b, (const float*)x_in, (float*)x_out, (float*)vel, box, mass, kick_dt, n_kicks, drift_dt, f_out, out12);
        return 0;
    }
    if (b->f32math && dtype == 0 && !b->use_cta) {
        b->launches += out12 ? 2 : 1;
        return host_inner_step_f32(h, (const float*)x_in, (float*)x_out, (float*)vel, box, mass, kick_dt, n_kicks,
                                   drift_dt, f_out, out12);
    }
    const Vec3d bx = {box[0], box[1], box[2]};
    b->launches += out12 ? 2 : 1;
    if (dtype == 1) inner<double>(b, 7, (const double*)x_in, (double*)x_out, (double*)vel, bx, mass, kick_dt, n_kicks, drift_dt, f_out, out12);
    else inner<float>(b, 7, (const float*)x_in, (float*)x_out, (float*)vel, bx, mass, kick_dt, n_kicks, drift_dt, f_out, out12);
    return 0;
}

// ---- k-space arithmetic of the GPE electrostatics (csrc/gpe.cuh) on host memory -----------------------
#include "../../hymd_b200/csrc/gpe.cuh"

extern "C" int host_gpe_kspace(const double* in, double* out_s, double* out_v, const double* tab, int Nx, int Ny,
                               int Nz, int F, double coef, int use_h, int div_k2, double sign) {
    GKParams p;
    p.Nx = Nx; p.Ny = Ny; p.Nz = Nz; p.nyl = Ny; p.y0 = 0; p.Nzc = Nz / 2 + 1; p.Nzcp = (p.Nzc + 1) / 2 * 2; p.F = F;
    const long long xs = (long long)Ny * p.Nzcp, fs = (long long)Nx * xs;
    p.npairs = fs / 2;
    p.xs_in = p.xs_s = p.xs_v = 2 * xs;
    p.fs_in = p.fs_s = p.fs_v = 2 * fs;
    for (long long i = 0; i < p.npairs; ++i)
        gpe_kspace_pair<double>(i, in, out_s, out_v, tab, coef, use_h, div_k2, sign, p);
    return 0;
}

// ---- single-precision bond / angle evaluators (csrc/bonded_f32.cuh) on host memory -------------------
#include "../../hymd_b200/csrc/bonded_f32.cuh"

extern "C" int host_inner_step_f32(void* h, const float* x_in, float* x_out, float* vel, const double* box,
                                   double mass, double kick_dt, int n_kicks, double drift_dt, void* const* f_out,
                                   double* out12) {
    HostBonded* b = (HostBonded*)h;
    TermLists t;
    for (int k = 0; k < 3; ++k) {
        t.start[k] = b->start[k].data(); t.refs[k] = b->refs[k].data(); t.idx[k] = b->idx[k].data();
        t.par[k] = b->par[k].data(); t.n_terms[k] = (long long)(b->idx[k].size() / 4);
    }
    t.dih_type = b->dtype.data();
    const Vec3d bx = {box[0], box[1], box[2]};
    float* fo[3] = {f_out ? (float*)f_out[0] : nullptr, f_out ? (float*)f_out[1] : nullptr,
                    f_out ? (float*)f_out[2] : nullptr};
    double acc12[12] = {0};
    for (long long p = 0; p < b->n; ++p) {
        BondAcc acc[3];
        inner_step_particle_f32(p, x_in, x_out, vel, bx, t, (float)mass, (float)(0.5 * kick_dt), n_kicks,
                                (float)drift_dt, fo, acc);
        for (int k = 0; k < 3; ++k) {
            acc12[4 * k] += acc[k].e; acc12[4 * k + 1] += acc[k].pr.x; acc12[4 * k + 2] += acc[k].pr.y;
            acc12[4 * k + 3] += acc[k].pr.z;
        }
    }
    if (out12) for (int k = 0; k < 12; ++k) out12[k] = acc12[k];
    return 0;
}

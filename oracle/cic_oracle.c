/* C restatement of pmesh's cloud-in-cell paint / readout loops.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  pmesh's own C sources are
 * a third-party dependency that is not vendored in /root/reference; this file
 * restates the published CIC algorithm as used at hymd/field.py:363, 574
 * (paint) and hymd/field.py:200, 402 (readout): vertex-centred mesh,
 * x = r*N/L, c = floor(x), d = x - c, weights (1-d | d) per axis, periodic
 * wrap of the vertex index, accumulation in particle order in the field dtype.
 *
 * The *_mt_* variants are the same arithmetic spread over POSIX threads (ORACLE_THREADS or all online cores) (the
 * stand-in for pmesh's one-MPI-rank-per-core execution); they are what
 * bench.py times as the CPU baseline.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static inline long wrap(long i, long n) { i %= n; return i < 0 ? i + n : i; }

static int n_threads(void) {
    const char *e = getenv("ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}

typedef struct {
    const void *a, *b; void *out, *priv;
    long n, m; int nx, ny, nz, t, nt; double lx, ly, lz;
    pthread_barrier_t *bar;
} job_t;

static void run_threads(void *(*fn)(void *), job_t *proto) {
    int nt = proto->nt;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * nt);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, nt);
    for (int t = 0; t < nt; ++t) { jobs[t] = *proto; jobs[t].t = t; jobs[t].bar = &bar; }
    for (int t = 1; t < nt; ++t) pthread_create(&th[t], NULL, fn, &jobs[t]);
    fn(&jobs[0]);
    for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&bar);
    free(th); free(jobs);
}

#define DEFINE_CIC(T, SUF, FLOOR)                                                          \
static inline void paint_one_##SUF(const T *p, T m, int nx, int ny, int nz, T sx, T sy,    \
                                   T sz, T *out) {                                         \
    T x = p[0] * sx, y = p[1] * sy, z = p[2] * sz;                                         \
    T fx = FLOOR(x), fy = FLOOR(y), fz = FLOOR(z);                                         \
    T dx = x - fx, dy = y - fy, dz = z - fz;                                               \
    long i0 = wrap((long)fx, nx), j0 = wrap((long)fy, ny), k0 = wrap((long)fz, nz);        \
    long i1 = i0 + 1 == nx ? 0 : i0 + 1, j1 = j0 + 1 == ny ? 0 : j0 + 1;                   \
    long k1 = k0 + 1 == nz ? 0 : k0 + 1;                                                   \
    T wx0 = (T)1 - dx, wy0 = (T)1 - dy, wz0 = (T)1 - dz;                                   \
    out[(i0 * ny + j0) * nz + k0] += m * wx0 * wy0 * wz0;                                  \
    out[(i0 * ny + j0) * nz + k1] += m * wx0 * wy0 * dz;                                   \
    out[(i0 * ny + j1) * nz + k0] += m * wx0 * dy * wz0;                                   \
    out[(i0 * ny + j1) * nz + k1] += m * wx0 * dy * dz;                                    \
    out[(i1 * ny + j0) * nz + k0] += m * dx * wy0 * wz0;                                   \
    out[(i1 * ny + j0) * nz + k1] += m * dx * wy0 * dz;                                    \
    out[(i1 * ny + j1) * nz + k0] += m * dx * dy * wz0;                                    \
    out[(i1 * ny + j1) * nz + k1] += m * dx * dy * dz;                                     \
}                                                                                          \
void cic_paint_##SUF(const T *pos, const T *mass, long n, int nx, int ny, int nz,          \
                     double lx, double ly, double lz, T *out) {                            \
    T sx = (T)(nx / lx), sy = (T)(ny / ly), sz = (T)(nz / lz);                             \
    for (long p = 0; p < n; ++p)                                                           \
        paint_one_##SUF(pos + 3 * p, mass[p], nx, ny, nz, sx, sy, sz, out);                \
}                                                                                          \
static void *paint_worker_##SUF(void *arg) {                                               \
    job_t *j = (job_t *)arg;                                                               \
    const T *pos = (const T *)j->a, *mass = (const T *)j->b;                               \
    T *out = (T *)j->out, *priv = (T *)j->priv;                                            \
    T sx = (T)(j->nx / j->lx), sy = (T)(j->ny / j->ly), sz = (T)(j->nz / j->lz);           \
    T *dst = j->t == 0 ? out : priv + (size_t)(j->t - 1) * j->m;                           \
    long lo = j->n * j->t / j->nt, hi = j->n * (j->t + 1) / j->nt;                         \
    for (long p = lo; p < hi; ++p)                                                         \
        paint_one_##SUF(pos + 3 * p, mass[p], j->nx, j->ny, j->nz, sx, sy, sz, dst);       \
    pthread_barrier_wait(j->bar);                                                          \
    long clo = j->m * j->t / j->nt, chi = j->m * (j->t + 1) / j->nt;                       \
    for (long c = clo; c < chi; ++c) {                                                     \
        T s = out[c];                                                                      \
        for (int u = 1; u < j->nt; ++u) s += priv[(size_t)(u - 1) * j->m + c];             \
        out[c] = s;                                                                        \
    }                                                                                      \
    return NULL;                                                                           \
}                                                                                          \
void cic_paint_mt_##SUF(const T *pos, const T *mass, long n, int nx, int ny, int nz,       \
                        double lx, double ly, double lz, T *out) {                         \
    job_t j; memset(&j, 0, sizeof j);                                                      \
    j.a = pos; j.b = mass; j.out = out; j.n = n; j.m = (long)nx * ny * nz;                 \
    j.nx = nx; j.ny = ny; j.nz = nz; j.lx = lx; j.ly = ly; j.lz = lz; j.nt = n_threads();  \
    j.priv = calloc((size_t)j.m * (j.nt - 1) + 1, sizeof(T));                              \
    run_threads(paint_worker_##SUF, &j);                                                   \
    free(j.priv);                                                                          \
}                                                                                          \
static inline T readout_one_##SUF(const T *f, const T *p, int nx, int ny, int nz, T sx,    \
                                  T sy, T sz) {                                            \
    T x = p[0] * sx, y = p[1] * sy, z = p[2] * sz;                                         \
    T fx = FLOOR(x), fy = FLOOR(y), fz = FLOOR(z);                                         \
    T dx = x - fx, dy = y - fy, dz = z - fz;                                               \
    long i0 = wrap((long)fx, nx), j0 = wrap((long)fy, ny), k0 = wrap((long)fz, nz);        \
    long i1 = i0 + 1 == nx ? 0 : i0 + 1, j1 = j0 + 1 == ny ? 0 : j0 + 1;                   \
    long k1 = k0 + 1 == nz ? 0 : k0 + 1;                                                   \
    T wx0 = (T)1 - dx, wy0 = (T)1 - dy, wz0 = (T)1 - dz;                                   \
    T v = 0;                                                                               \
    v += wx0 * wy0 * wz0 * f[(i0 * ny + j0) * nz + k0];                                    \
    v += wx0 * wy0 * dz * f[(i0 * ny + j0) * nz + k1];                                     \
    v += wx0 * dy * wz0 * f[(i0 * ny + j1) * nz + k0];                                     \
    v += wx0 * dy * dz * f[(i0 * ny + j1) * nz + k1];                                      \
    v += dx * wy0 * wz0 * f[(i1 * ny + j0) * nz + k0];                                     \
    v += dx * wy0 * dz * f[(i1 * ny + j0) * nz + k1];                                      \
    v += dx * dy * wz0 * f[(i1 * ny + j1) * nz + k0];                                      \
    v += dx * dy * dz * f[(i1 * ny + j1) * nz + k1];                                       \
    return v;                                                                              \
}                                                                                          \
void cic_readout_##SUF(const T *f, const T *pos, long n, int nx, int ny, int nz,           \
                       double lx, double ly, double lz, T *out) {                          \
    T sx = (T)(nx / lx), sy = (T)(ny / ly), sz = (T)(nz / lz);                             \
    for (long p = 0; p < n; ++p)                                                           \
        out[p] = readout_one_##SUF(f, pos + 3 * p, nx, ny, nz, sx, sy, sz);                \
}                                                                                          \
static void *readout_worker_##SUF(void *arg) {                                             \
    job_t *j = (job_t *)arg;                                                               \
    const T *f = (const T *)j->a, *pos = (const T *)j->b;                                  \
    T *out = (T *)j->out;                                                                  \
    T sx = (T)(j->nx / j->lx), sy = (T)(j->ny / j->ly), sz = (T)(j->nz / j->lz);           \
    long lo = j->n * j->t / j->nt, hi = j->n * (j->t + 1) / j->nt;                         \
    for (long p = lo; p < hi; ++p)                                                         \
        out[p] = readout_one_##SUF(f, pos + 3 * p, j->nx, j->ny, j->nz, sx, sy, sz);       \
    return NULL;                                                                           \
}                                                                                          \
void cic_readout_mt_##SUF(const T *f, const T *pos, long n, int nx, int ny, int nz,        \
                          double lx, double ly, double lz, T *out) {                       \
    job_t j; memset(&j, 0, sizeof j);                                                      \
    j.a = f; j.b = pos; j.out = out; j.n = n;                                              \
    j.nx = nx; j.ny = ny; j.nz = nz; j.lx = lx; j.ly = ly; j.lz = lz; j.nt = n_threads();  \
    run_threads(readout_worker_##SUF, &j);                                                 \
}

DEFINE_CIC(float, f32, floorf)
DEFINE_CIC(double, f64, floor)

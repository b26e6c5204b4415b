/*
 * oracle/g3d_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the reference's distance-field
 * path (the vendored SDFGen core + the dfgen driver arithmetic + the vox2mesh
 * occupancy rule). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker or the CPU arm -- the product path (generate3d_b200/) never does.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md section 4/8c), so this restatement is pinned against the
 * reference's own sources compiled in place (oracle/_ref/libg3d_ref.so, see
 * oracle/Makefile + oracle/ref_shim.cpp); tests/test_oracle.py compares
 * the two bit-for-bit whenever oracle/_ref exists.
 *
 * Build: gcc -O2 -ffp-contract=off (no -march, no fast-math) so that every
 * float operation is a separately rounded IEEE binary32 op, exactly like the
 * reference binary ("c++ -std=c++14 -O3", src/dfgen/makefile:3-4, baseline
 * x86-64 => no FMA).
 *
 * Every function cites the reference file:line it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* std::min / std::max comparison direction matters for NaN (util.h:18-20). */
static inline float min_f(float a, float b) { return (b < a) ? b : a; }
static inline float max_f(float a, float b) { return (a < b) ? b : a; }
static inline double min_d(double a, double b) { return (b < a) ? b : a; }
static inline double max_d(double a, double b) { return (a < b) ? b : a; }
/* util.h:30-32 and :46-48: three-argument forms nest on the right. */
static inline double min3_d(double a, double b, double c) { return min_d(a, min_d(b, c)); }
static inline double max3_d(double a, double b, double c) { return max_d(a, max_d(b, c)); }
/* util.h:169-175 */
static inline int clamp_i(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }

/* vec.h:327-333 dot, :198-204 mag2, :210-220 dist: left-to-right float sums. */
static inline float dot3(const float *a, const float *b)
{
    float d = a[0] * b[0];
    d += a[1] * b[1];
    d += a[2] * b[2];
    return d;
}

static inline float dist3(const float *a, const float *b)
{
    float t0 = a[0] - b[0], t1 = a[1] - b[1], t2 = a[2] - b[2];
    float d = t0 * t0;
    d += t1 * t1;
    d += t2 * t2;
    return sqrtf(d);
}

/* makelevelset3.cpp:4-17 */
float g3d_oracle_point_segment_distance(const float *x0, const float *x1, const float *x2)
{
    float e[3] = { x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2] };
    double m2 = (double)dot3(e, e);                 /* float mag2 widened (:7) */
    float r[3] = { x2[0] - x0[0], x2[1] - x0[1], x2[2] - x0[2] };
    float s12 = (float)((double)dot3(r, e) / m2);   /* double division (:9) */
    if (s12 < 0) s12 = 0;
    else if (s12 > 1) s12 = 1;
    float one_minus = 1.0f - s12;
    float c[3];
    for (int i = 0; i < 3; ++i) c[i] = x1[i] * s12 + x2[i] * one_minus;
    return dist3(x0, c);
}

/* makelevelset3.cpp:20-41 */
float g3d_oracle_point_triangle_distance(const float *x0, const float *x1, const float *x2,
                                         const float *x3)
{
    float x13[3], x23[3], x03[3];
    for (int i = 0; i < 3; ++i) {
        x13[i] = x1[i] - x3[i];
        x23[i] = x2[i] - x3[i];
        x03[i] = x0[i] - x3[i];
    }
    float m13 = dot3(x13, x13), m23 = dot3(x23, x23), d = dot3(x13, x23);
    float invdet = 1.f / max_f(m13 * m23 - d * d, 1e-30f);
    float a = dot3(x13, x03), b = dot3(x23, x03);
    float w23 = invdet * (m23 * a - d * b);
    float w31 = invdet * (m13 * b - d * a);
    float w12 = 1 - w23 - w31;
    if (w23 >= 0 && w31 >= 0 && w12 >= 0) {
        float c[3];
        for (int i = 0; i < 3; ++i) c[i] = x1[i] * w23 + x2[i] * w31 + x3[i] * w12;
        return dist3(x0, c);
    }
    if (w23 > 0)
        return min_f(g3d_oracle_point_segment_distance(x0, x1, x2),
                     g3d_oracle_point_segment_distance(x0, x1, x3));
    if (w31 > 0)
        return min_f(g3d_oracle_point_segment_distance(x0, x1, x2),
                     g3d_oracle_point_segment_distance(x0, x2, x3));
    return min_f(g3d_oracle_point_segment_distance(x0, x1, x3),
                 g3d_oracle_point_segment_distance(x0, x2, x3));
}

/* makelevelset3.cpp:84-94 */
static int orientation2d(double x1, double y1, double x2, double y2, double *area2)
{
    *area2 = y1 * x2 - x1 * y2;
    if (*area2 > 0) return 1;
    if (*area2 < 0) return -1;
    if (y2 > y1) return 1;
    if (y2 < y1) return -1;
    if (x1 > x2) return 1;
    if (x1 < x2) return -1;
    return 0;
}

/* makelevelset3.cpp:98-116; returns 1 when inside and sets barycentrics. */
static int point_in_triangle_2d(double x0, double y0, double x1, double y1, double x2, double y2,
                                double x3, double y3, double *a, double *b, double *c)
{
    x1 -= x0; x2 -= x0; x3 -= x0;
    y1 -= y0; y2 -= y0; y3 -= y0;
    int sa = orientation2d(x2, y2, x3, y3, a);
    if (sa == 0) return 0;
    int sb = orientation2d(x3, y3, x1, y1, b);
    if (sb != sa) return 0;
    int sc = orientation2d(x1, y1, x2, y2, c);
    if (sc != sa) return 0;
    double sum = *a + *b + *c;
    *a /= sum; *b /= sum; *c /= sum;
    return 1;
}

typedef struct {
    const uint32_t *tri;
    const float *x;
    float *phi;
    int32_t *ct;
    int ni, nj, nk;
    float origin[3];
    float dx;
} ls_state;

#define IDX(s, i, j, k) ((size_t)(i) + (size_t)(s)->ni * ((size_t)(j) + (size_t)(s)->nj * (size_t)(k)))

/* makelevelset3.cpp:43-55 */
static inline void check_neighbour(ls_state *s, const float *gx, size_t n0, size_t n1)
{
    int32_t t = s->ct[n1];
    if (t >= 0) {
        const uint32_t *v = s->tri + 3 * (size_t)t;
        float d = g3d_oracle_point_triangle_distance(gx, s->x + 3 * (size_t)v[0],
                                                     s->x + 3 * (size_t)v[1],
                                                     s->x + 3 * (size_t)v[2]);
        if (d < s->phi[n0]) {
            s->phi[n0] = d;
            s->ct[n0] = t;
        }
    }
}

/* makelevelset3.cpp:57-80 */
static void sweep(ls_state *s, int di, int dj, int dk)
{
    int i0 = di > 0 ? 1 : s->ni - 2, i1 = di > 0 ? s->ni : -1;
    int j0 = dj > 0 ? 1 : s->nj - 2, j1 = dj > 0 ? s->nj : -1;
    int k0 = dk > 0 ? 1 : s->nk - 2, k1 = dk > 0 ? s->nk : -1;
    for (int k = k0; k != k1; k += dk)
        for (int j = j0; j != j1; j += dj)
            for (int i = i0; i != i1; i += di) {
                float gx[3] = { i * s->dx + s->origin[0], j * s->dx + s->origin[1],
                                k * s->dx + s->origin[2] };
                size_t n = IDX(s, i, j, k);
                check_neighbour(s, gx, n, IDX(s, i - di, j, k));
                check_neighbour(s, gx, n, IDX(s, i, j - dj, k));
                check_neighbour(s, gx, n, IDX(s, i - di, j - dj, k));
                check_neighbour(s, gx, n, IDX(s, i, j, k - dk));
                check_neighbour(s, gx, n, IDX(s, i - di, j, k - dk));
                check_neighbour(s, gx, n, IDX(s, i, j - dj, k - dk));
                check_neighbour(s, gx, n, IDX(s, i - di, j - dj, k - dk));
            }
}

/*
 * makelevelset3.cpp:118-183. `stage` lets unit-level parity tests stop early:
 *   0 = full function, 1 = stop after the band init + intersection counts,
 *   2 = stop after the 16 sweeps (unsigned distances).
 * closest_tri_out / isect_out may be NULL. Returns the number of
 * point-triangle evaluations performed (for the work accounting in bench.py).
 */
int64_t g3d_oracle_make_level_set3(const uint32_t *tri, int64_t ntri, const float *x,
                                   const float *origin, float dx, int ni, int nj, int nk,
                                   float *phi, int32_t *closest_tri_out, int32_t *isect_out,
                                   int exact_band, int stage)
{
    size_t n = (size_t)ni * nj * nk;
    ls_state s;
    s.tri = tri; s.x = x; s.phi = phi; s.ni = ni; s.nj = nj; s.nk = nk; s.dx = dx;
    memcpy(s.origin, origin, sizeof s.origin);
    s.ct = (int32_t *)malloc(n * sizeof(int32_t));
    int32_t *cnt = (int32_t *)calloc(n, sizeof(int32_t));
    int64_t evals = 0;
    float far = (ni + nj + nk) * dx;                               /* :123 */
    for (size_t q = 0; q < n; ++q) { phi[q] = far; s.ct[q] = -1; }

    for (int64_t t = 0; t < ntri; ++t) {                           /* :128 */
        const float *P = x + 3 * (size_t)tri[3 * t], *Q = x + 3 * (size_t)tri[3 * t + 1],
                    *R = x + 3 * (size_t)tri[3 * t + 2];
        double fp[3], fq[3], fr[3];                                /* :131-133 */
        for (int a = 0; a < 3; ++a) {
            fp[a] = ((double)P[a] - origin[a]) / dx;
            fq[a] = ((double)Q[a] - origin[a]) / dx;
            fr[a] = ((double)R[a] - origin[a]) / dx;
        }
        int dims[3] = { ni, nj, nk }, lo[3], hi[3];                /* :135-137 */
        for (int a = 0; a < 3; ++a) {
            lo[a] = clamp_i((int)min3_d(fp[a], fq[a], fr[a]) - exact_band, 0, dims[a] - 1);
            hi[a] = clamp_i((int)max3_d(fp[a], fq[a], fr[a]) + exact_band + 1, 0, dims[a] - 1);
        }
        for (int k = lo[2]; k <= hi[2]; ++k)                       /* :138-146 */
            for (int j = lo[1]; j <= hi[1]; ++j)
                for (int i = lo[0]; i <= hi[0]; ++i) {
                    float gx[3] = { i * dx + origin[0], j * dx + origin[1], k * dx + origin[2] };
                    float d = g3d_oracle_point_triangle_distance(gx, P, Q, R);
                    ++evals;
                    size_t q = IDX(&s, i, j, k);
                    if (d < phi[q]) { phi[q] = d; s.ct[q] = (int32_t)t; }
                }
        /* :147-160 x-ray intersection counts */
        int j0 = clamp_i((int)ceil(min3_d(fp[1], fq[1], fr[1])), 0, nj - 1);
        int j1 = clamp_i((int)floor(max3_d(fp[1], fq[1], fr[1])), 0, nj - 1);
        int k0 = clamp_i((int)ceil(min3_d(fp[2], fq[2], fr[2])), 0, nk - 1);
        int k1 = clamp_i((int)floor(max3_d(fp[2], fq[2], fr[2])), 0, nk - 1);
        for (int k = k0; k <= k1; ++k)
            for (int j = j0; j <= j1; ++j) {
                double a, b, c;
                if (point_in_triangle_2d(j, k, fp[1], fp[2], fq[1], fq[2], fr[1], fr[2],
                                         &a, &b, &c)) {
                    double fi = a * fp[0] + b * fq[0] + c * fr[0];
                    int iv = (int)ceil(fi);
                    if (iv < 0) ++cnt[IDX(&s, 0, j, k)];
                    else if (iv < ni) ++cnt[IDX(&s, iv, j, k)];
                }
            }
    }
    if (stage != 1) {
        /* :163-172; count evaluations = neighbours with a triangle */
        static const int dirs[8][3] = { { +1, +1, +1 }, { -1, -1, -1 }, { +1, +1, -1 },
                                        { -1, -1, +1 }, { +1, -1, +1 }, { -1, +1, -1 },
                                        { +1, -1, -1 }, { -1, +1, +1 } };
        for (int pass = 0; pass < 2; ++pass)
            for (int q = 0; q < 8; ++q) sweep(&s, dirs[q][0], dirs[q][1], dirs[q][2]);
        evals += (int64_t)16 * 7 * (ni - 1) * (nj - 1) * (nk - 1); /* upper bound */
    }
    if (stage == 0) {
        for (int k = 0; k < nk; ++k)                               /* :174-182 */
            for (int j = 0; j < nj; ++j) {
                int total = 0;
                for (int i = 0; i < ni; ++i) {
                    size_t q = IDX(&s, i, j, k);
                    total += cnt[q];
                    if (total % 2 == 1) phi[q] = -phi[q];
                }
            }
    }
    if (closest_tri_out) memcpy(closest_tri_out, s.ct, n * sizeof(int32_t));
    if (isect_out) memcpy(isect_out, cnt, n * sizeof(int32_t));
    free(s.ct);
    free(cnt);
    return evals;
}

/*
 * dfgen driver arithmetic, src/dfgen/main.cpp:45-121, for the only geometry the
 * shipped pipeline uses or that passes its assert (:102). Fills origin/dx/sizes
 * and the row-major grid2world written to the .df header (:106-114).
 * Returns 0, or -1 when the reference would trip assert(dims == dim).
 */
int g3d_oracle_dfgen_geometry(int dim, int padding, int is_unit_cube, const float *bbox_min,
                              const float *bbox_max, float *origin, float *dx_out, int *sizes,
                              float *grid2world)
{
    float dx = 1.0 / (dim - 2 * padding);                          /* :57 double -> float */
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        lo[a] = is_unit_cube ? (-1.0f * 0.5f) : bbox_min[a];      /* :82-85 */
        hi[a] = is_unit_cube ? (1.0f * 0.5f) : bbox_max[a];
        float pad = padding * dx * 1.0f;                           /* :91-93 (padding*dx)*unit */
        lo[a] -= pad;
        hi[a] += pad;
        sizes[a] = (int)(unsigned int)((hi[a] - lo[a]) / dx);      /* :94 */
        origin[a] = lo[a];
    }
    *dx_out = dx;
    float dfix = 1.0 / dim;                                        /* :107 */
    float tr = padding != 0 ? (-0.5f * 1.0f - dfix) : (-0.5f * 1.0f);
    memset(grid2world, 0, 16 * sizeof(float));
    grid2world[0] = grid2world[5] = grid2world[10] = 1.0f * dx;    /* :109 */
    grid2world[3] = grid2world[7] = grid2world[11] = tr;           /* :110-113 */
    grid2world[15] = 1.0f;
    return (sizes[0] == dim && sizes[1] == dim && sizes[2] == dim) ? 0 : -1;
}

/* main.cpp:117-120: voxel-unit scaling of the signed distances. */
void g3d_oracle_scale_sdf(const float *phi, int64_t n, int dim, float *sdf)
{
    for (int64_t i = 0; i < n; ++i) sdf[i] = phi[i] * dim;
}

/* src/utils/LoaderVOX.h:48-63 (.df writer; no pdf block). */
int g3d_oracle_write_df(const char *path, const int32_t *dims, float res, const float *grid2world,
                        const float *sdf)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    fwrite(dims, sizeof(int32_t), 3, f);
    fwrite(&res, sizeof(float), 1, f);
    fwrite(grid2world, sizeof(float), 16, f);
    fwrite(sdf, sizeof(float), (size_t)dims[0] * dims[1] * dims[2], f);
    fclose(f);
    return 0;
}

/* src/vox2mesh/main.cpp:86-104: occupancy rule |sdf| <= trunc*res. */
void g3d_oracle_occupancy(const float *sdf, int64_t n, float trunc, float res, uint8_t *occ)
{
    float thr = trunc * res;
    for (int64_t i = 0; i < n; ++i) occ[i] = fabsf(sdf[i]) <= thr;
}

/*
 * Brute force: for every node the minimum of the reference's distance function
 * (makelevelset3.cpp:20-41) over ALL triangles, lowest triangle index on ties.
 * This is what make_level_set3 would return if its band covered the whole grid;
 * it is the checker for the product's "exact" mode.
 */
void g3d_oracle_brute_force(const uint32_t *tri, int64_t ntri, const float *x, const float *origin,
                            float dx, int ni, int nj, int nk, float *phi, int32_t *closest_tri)
{
    float far = (ni + nj + nk) * dx;
    for (int k = 0; k < nk; ++k)
        for (int j = 0; j < nj; ++j)
            for (int i = 0; i < ni; ++i) {
                float gx[3] = { i * dx + origin[0], j * dx + origin[1], k * dx + origin[2] };
                float best = far;
                int32_t bt = -1;
                for (int64_t t = 0; t < ntri; ++t) {
                    float d = g3d_oracle_point_triangle_distance(gx, x + 3 * (size_t)tri[3 * t],
                                                                 x + 3 * (size_t)tri[3 * t + 1],
                                                                 x + 3 * (size_t)tri[3 * t + 2]);
                    if (d < best) { best = d; bt = (int32_t)t; }
                }
                size_t q = (size_t)i + (size_t)ni * ((size_t)j + (size_t)nj * (size_t)k);
                phi[q] = best;
                if (closest_tri) closest_tri[q] = bt;
            }
}

/* The same minimum for a list of voxels only (linear indices i + ni*(j + nj*k)): lets the GPU tests spot-check grids
 * whose full brute force would take hours on a CPU (128^3 x 200k triangles). */
void g3d_oracle_brute_force_points(const uint32_t *tri, int64_t ntri, const float *x, const float *origin,
                                   float dx, int ni, int nj, int nk, const int64_t *idx, int64_t nidx,
                                   float *phi, int32_t *closest_tri)
{
    float far = (ni + nj + nk) * dx;
    for (int64_t p = 0; p < nidx; ++p) {
        int64_t q = idx[p];
        int i = (int)(q % ni), j = (int)((q / ni) % nj), k = (int)(q / ((int64_t)ni * nj));
        float gx[3] = { i * dx + origin[0], j * dx + origin[1], k * dx + origin[2] };
        float best = far;
        int32_t bt = -1;
        for (int64_t t = 0; t < ntri; ++t) {
            float d = g3d_oracle_point_triangle_distance(gx, x + 3 * (size_t)tri[3 * t],
                                                         x + 3 * (size_t)tri[3 * t + 1],
                                                         x + 3 * (size_t)tri[3 * t + 2]);
            if (d < best) { best = d; bt = (int32_t)t; }
        }
        phi[p] = best;
        if (closest_tri) closest_tri[p] = bt;
    }
}

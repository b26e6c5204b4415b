/* tools/sweep_stats.c -- DEVELOPMENT TOOL (not product, not oracle): replays the reference's 16 sweeps on the CPU with the
 * sweep kernels' bookkeeping and counts, per sweep, what each shortcut of generate3d_b200/csrc/g3d_tdf.cu leaves to do.
 * Built and driven by tools/sweep_stats.py. Includes the C oracle for the distance function and the band init.
 *
 * out[s][0..7]: node visits | nodes with a candidate under rules (1)-(2) (own / earlier candidate, unchanged neighbour) |
 *               evaluations under (1)-(2) | of those: pair evaluated in an earlier sweep | node inside the triangle's band box |
 *               evaluations left by (3) met neighbours | left by (3) + (4) band-box flags | nodes changed */
#include "../oracle/g3d_oracle.c"

static void sdir(int s, int *d)
{
    int q8 = s & 7;
    d[0] = (q8 & 1) ? -1 : 1;
    d[1] = (q8 == 0 || q8 == 2 || q8 == 5 || q8 == 7) ? 1 : -1;
    d[2] = (q8 == 0 || q8 == 3 || q8 == 4 || q8 == 7) ? 1 : -1;
}

static int threshold(int s, int mask)         /* code of the last sweep before s that ran like s on the axes of mask */
{
    int d[3], thr = -1;
    sdir(s, d);
    for (int sp = 0; sp < s; ++sp) {
        int e[3], ok = 1;
        sdir(sp, e);
        for (int ax = 0; ax < 3; ++ax)
            if (((mask >> ax) & 1) && e[ax] != d[ax]) ok = 0;
        if (ok) thr = sp + 1;
    }
    return thr;
}

#define HIST 64                               /* triangles remembered per node for the "evaluated before" count */

void sweep_stats(const uint32_t *tri, int64_t ntri, const float *x, const float *origin, float dx, int ni, int nj, int nk,
                 int64_t *out)
{
    size_t n = (size_t)ni * nj * nk;
    float *phi = (float *)malloc(n * 4);
    int32_t *ct = (int32_t *)malloc(n * 4);
    g3d_oracle_make_level_set3(tri, ntri, x, origin, dx, ni, nj, nk, phi, ct, NULL, 1, 1);       /* band init only */
    int8_t *code = (int8_t *)calloc(n, 1);
    int32_t *hist = (int32_t *)malloc(n * HIST * 4);
    int *hn = (int *)calloc(n, 4);
    int dims[3] = { ni, nj, nk };
    memset(out, 0, 16 * 8 * sizeof(int64_t));
    for (int s = 0; s < 16; ++s) {
        int d[3];
        sdir(s, d);
        int64_t *o = out + 8 * s;
        /* raster order = the reference's; the counts do not depend on the order inside a level */
        for (int c = 0; c < nk - 1; ++c) for (int b = 0; b < nj - 1; ++b) for (int a = 0; a < ni - 1; ++a) {
            int p[3] = { d[0] > 0 ? 1 + a : ni - 2 - a, d[1] > 0 ? 1 + b : nj - 2 - b, d[2] > 0 ? 1 + c : nk - 2 - c };
            size_t q = (size_t)p[0] + (size_t)ni * (p[1] + (size_t)nj * p[2]);
            int face = 0;
            for (int ax = 0; ax < 3; ++ax) if (p[ax] == 0 || p[ax] == dims[ax] - 1) face |= 1 << ax;
            int32_t t[7]; int unchanged[7], inband[7];
            for (int r = 0; r < 7; ++r) {
                int m[3] = { p[0] - (((r + 1) & 1) ? d[0] : 0), p[1] - (((r + 1) & 2) ? d[1] : 0), p[2] - (((r + 1) & 4) ? d[2] : 0) };
                size_t qm = (size_t)m[0] + (size_t)ni * (m[1] + (size_t)nj * m[2]);
                t[r] = ct[qm];
                unchanged[r] = code[qm] <= threshold(s, (r + 1) | face);
                inband[r] = 0;
                if (t[r] >= 0 && code[qm] == 0) {        /* still the band-init triangle: is this node in its band box? */
                    const uint32_t *v = tri + 3 * (size_t)t[r];
                    inband[r] = 1;
                    for (int ax = 0; ax < 3; ++ax) {
                        double f0 = ((double)x[3 * v[0] + ax] - origin[ax]) / dx, f1 = ((double)x[3 * v[1] + ax] - origin[ax]) / dx,
                               f2 = ((double)x[3 * v[2] + ax] - origin[ax]) / dx;
                        int lo = clamp_i((int)min3_d(f0, f1, f2) - 1, 0, dims[ax] - 1), hi = clamp_i((int)max3_d(f0, f1, f2) + 2, 0, dims[ax] - 1);
                        if (p[ax] < lo || p[ax] > hi) inband[r] = 0;
                    }
                }
            }
            float gx[3] = { p[0] * dx + origin[0], p[1] * dx + origin[1], p[2] * dx + origin[2] };
            const int32_t own = ct[q];
            int any = 0, changed = 0;
            ++o[0];
            for (int r = 0; r < 7; ++r) {
                if (t[r] < 0 || unchanged[r]) continue;
                any = 1;
                int dup = t[r] == own;
                for (int z = 0; z < r; ++z) if (t[z] == t[r] && t[z] >= 0 && !unchanged[z]) dup = 1;   /* rules (1)-(2) */
                if (dup) continue;
                ++o[2];
                int seen = 0;
                for (int z = 0; z < hn[q]; ++z) if (hist[q * HIST + z] == t[r]) seen = 1;
                if (!seen && hn[q] < HIST) hist[q * HIST + hn[q]++] = t[r];
                o[3] += seen;
                o[4] += inband[r];
                int met3 = 0, met4 = inband[r];
                for (int z = 0; z < 7; ++z) {
                    if (z == r || t[z] != t[r]) continue;
                    if (unchanged[z]) met3 = 1;                                  /* rule (3) */
                    if (unchanged[z] || inband[z]) met4 = 1;                     /* rules (3) + (4) */
                }
                o[5] += !met3;
                o[6] += !met4;
                const uint32_t *v = tri + 3 * (size_t)t[r];
                float dd = g3d_oracle_point_triangle_distance(gx, x + 3 * (size_t)v[0], x + 3 * (size_t)v[1], x + 3 * (size_t)v[2]);
                if (dd < phi[q]) { phi[q] = dd; ct[q] = t[r]; changed = 1; }
            }
            o[1] += any;
            if (changed) { code[q] = (int8_t)(s + 1); ++o[7]; }
        }
    }
    free(phi); free(ct); free(code); free(hist); free(hn);
}

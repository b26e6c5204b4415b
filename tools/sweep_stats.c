/* tools/sweep_stats.c -- DEVELOPMENT TOOL (not product, not oracle): replays the reference's 16 sweeps on the CPU with
 * the sweep kernel's bookkeeping (change stamps, re-examination thresholds) and counts what finer-grained skipping
 * could save. Built and driven by tools/sweep_stats.py. Includes the C oracle for the distance function. */
#include "../oracle/g3d_oracle.c"

static void sdir(int s, int *d)
{
    int q8 = s & 7;
    d[0] = (q8 & 1) ? -1 : 1;
    d[1] = (q8 == 0 || q8 == 2 || q8 == 5 || q8 == 7) ? 1 : -1;
    d[2] = (q8 == 0 || q8 == 3 || q8 == 4 || q8 == 7) ? 1 : -1;
}

/* out[s][0..7]: visits, stamp-active nodes, nodes with >= 1 evaluation, evaluations, changes,
 *               32-node rounds in all levels, rounds holding a stamp-active node, levels holding one */
void sweep_stats(const uint32_t *tri, int64_t ntri, const float *x, const float *origin, float dx, int ni, int nj,
                 int nk, int64_t *out, int64_t *extra, int64_t *extra2, int64_t *extra3)
{
    /* extra[s][0]: evals with n inside band box of t; [1]: evals of a pair evaluated in an earlier sweep; [2]: either */
    uint64_t *hist = (uint64_t *)calloc((size_t)ni * nj * nk * 64, 8); int *hn = (int *)calloc((size_t)ni * nj * nk, 4);
    size_t n = (size_t)ni * nj * nk;
    uint32_t *fifo = (uint32_t *)malloc(n * 8 * 3 * 4); memset(fifo, 0xff, n * 8 * 3 * 4); int *fp_ = (int *)calloc(n * 3, 4);
    float *phi = (float *)malloc(n * 4);
    int32_t *ct = (int32_t *)malloc(n * 4);
    g3d_oracle_make_level_set3(tri, ntri, x, origin, dx, ni, nj, nk, phi, ct, NULL, 1, 1);
    int8_t *stamp = (int8_t *)calloc(n, 1);
    /* last[s'][...] not needed: thresholds from direction history */
    int dims[3] = { ni, nj, nk };
    for (int s = 0; s < 16; ++s) {
        int d[3];
        sdir(s, d);
        int na = ni - 1, nb = nj - 1, nc = nk - 1, nlev = na + nb + nc - 2;
        int64_t *o = out + 8 * s;
        memset(o, 0, 64);
        /* rounds: nodes of a level in (c, b) order, 32 at a time */
        for (int L = 0; L < nlev; ++L) {
            int in_round = 0, round_active = 0, level_active = 0, cnt = 0;
            for (int c = 0; c < nc; ++c)
                for (int b = 0; b < nb; ++b) {
                    int a = L - b - c;
                    if (a < 0 || a >= na) continue;
                    int p[3] = { d[0] > 0 ? 1 + a : ni - 2 - a, d[1] > 0 ? 1 + b : nj - 2 - b, d[2] > 0 ? 1 + c : nk - 2 - c };
                    size_t q = (size_t)p[0] + (size_t)ni * (p[1] + (size_t)nj * p[2]);
                    ++o[0];
                    int cand[7], ncand = 0, active = 0, evals = 0, changed = 0;
                    float gx[3] = { p[0] * dx + origin[0], p[1] * dx + origin[1], p[2] * dx + origin[2] };
                    int ex_t[7], n_ex = 0;              /* triangles of upwind neighbours already examined and unchanged since */
                    for (int r = 1; r <= 7; ++r) {
                        int m[3] = { p[0] - ((r & 1) ? d[0] : 0), p[1] - ((r & 2) ? d[1] : 0), p[2] - ((r & 4) ? d[2] : 0) };
                        size_t qm = (size_t)m[0] + (size_t)ni * (m[1] + (size_t)nj * m[2]);
                        int thr = -1;
                        for (int sp = 0; sp < s; ++sp) {
                            int e[3], ok = 1;
                            sdir(sp, e);
                            for (int ax = 0; ax < 3; ++ax) {
                                if (((r >> ax) & 1) && e[ax] != d[ax]) ok = 0;
                                if (e[ax] > 0 ? p[ax] == 0 : p[ax] == dims[ax] - 1) ok = 0;
                            }
                            if (ok) thr = sp + 1;
                        }
                        if (ct[qm] >= 0 && stamp[qm] <= thr) ex_t[n_ex++] = ct[qm];
                    }
                    for (int r = 1; r <= 7; ++r) {
                        int m[3] = { p[0] - ((r & 1) ? d[0] : 0), p[1] - ((r & 2) ? d[1] : 0), p[2] - ((r & 4) ? d[2] : 0) };
                        size_t qm = (size_t)m[0] + (size_t)ni * (m[1] + (size_t)nj * m[2]);
                        /* threshold: last sweep sp < s with the same neighbour upwind and the node swept */
                        int thr = -1;
                        for (int sp = 0; sp < s; ++sp) {
                            int e[3], ok = 1;
                            sdir(sp, e);
                            for (int ax = 0; ax < 3; ++ax) {
                                if (((r >> ax) & 1) && e[ax] != d[ax]) ok = 0;
                                if (e[ax] > 0 ? p[ax] == 0 : p[ax] == dims[ax] - 1) ok = 0;
                            }
                            if (ok) thr = sp + 1;
                        }
                        int t = ct[qm];
                        if (t < 0 || stamp[qm] <= thr) continue;
                        active = 1;
                        int dup = (t == ct[q]);
                        for (int z = 0; z < ncand; ++z) if (cand[z] == t) dup = 1;
                        cand[ncand++] = t;
                        if (dup) continue;
                        ++evals;
                        { int exd = 0; for (int z = 0; z < n_ex; ++z) if (ex_t[z] == t) exd = 1; extra[4 * s + 3] += exd;
                          if (!exd) {
                            /* m = neighbour qm holds t; flags only valid if stamp[qm] == 0 (init-assigned) */
                            const uint32_t *vv = tri + 3 * (size_t)t; int inb = 1, interior = 1, dirok = 1;
                            for (int ax = 0; ax < 3; ++ax) {
                                double f0 = ((double)x[3*vv[0]+ax] - origin[ax]) / dx, f1 = ((double)x[3*vv[1]+ax] - origin[ax]) / dx, f2 = ((double)x[3*vv[2]+ax] - origin[ax]) / dx;
                                int lo = clamp_i((int)min3_d(f0, f1, f2) - 1, 0, dims[ax] - 1), hi = clamp_i((int)max3_d(f0, f1, f2) + 2, 0, dims[ax] - 1);
                                if (p[ax] < lo || p[ax] > hi) inb = 0;
                                if (!(m[ax] > lo && m[ax] < hi)) interior = 0;
                                if (p[ax] != m[ax] && (p[ax] < lo || p[ax] > hi)) dirok = 0;
                            }
                            int init_assigned = stamp[qm] == 0;
                            extra2[8 * s + 7] += 1;                                   /* evals left after dup rule */
                            extra[4 * s + 0] += 0;
                            extra3[4 * s + 0] += inb;                                  /* full in-band knowledge */
                            extra3[4 * s + 1] += (init_assigned && interior);          /* 1-bit flag */
                            extra3[4 * s + 2] += (init_assigned && dirok);             /* 6-bit flags */
                          }
                        }
                        {
                            const uint32_t *vv = tri + 3 * (size_t)t; int inb = 1;
                            for (int ax = 0; ax < 3; ++ax) {
                                double f0 = ((double)x[3*vv[0]+ax] - origin[ax]) / dx, f1 = ((double)x[3*vv[1]+ax] - origin[ax]) / dx, f2 = ((double)x[3*vv[2]+ax] - origin[ax]) / dx;
                                int lo = clamp_i((int)min3_d(f0, f1, f2) - 1, 0, dims[ax] - 1), hi = clamp_i((int)max3_d(f0, f1, f2) + 2, 0, dims[ax] - 1);
                                if (p[ax] < lo || p[ax] > hi) inb = 0;
                            }
                            int seen = 0;
                            for (int z = 0; z < hn[q]; ++z) if (hist[q * 64 + z] == (uint64_t)t) seen = 1;
                            if (!seen && hn[q] < 64) hist[q * 64 + hn[q]++] = (uint64_t)t;
                            {   /* FIFO caches of 2, 4, 8 entries; triangle inequality */
                                static const int ks[3] = { 2, 4, 8 };
                                int hitk[3] = { 0, 0, 0 };
                                for (int c3 = 0; c3 < 3; ++c3) {
                                    uint32_t *f = fifo + (q * 3 + c3) * 8;
                                    for (int z = 0; z < ks[c3]; ++z) if (f[z] == (uint32_t)t) hitk[c3] = 1;
                                    if (!hitk[c3]) { f[fp_[q * 3 + c3]] = (uint32_t)t; fp_[q * 3 + c3] = (fp_[q * 3 + c3] + 1) % ks[c3]; }
                                }
                                float step = dx * sqrtf((float)((r & 1) + ((r >> 1) & 1) + ((r >> 2) & 1)));
                                int ti = (phi[qm] - step * 1.0001f >= phi[q]);
                                extra2[8 * s + 0] += hitk[0]; extra2[8 * s + 1] += hitk[1]; extra2[8 * s + 2] += hitk[2];
                                extra2[8 * s + 3] += ti; extra2[8 * s + 4] += (ti || hitk[1]); extra2[8 * s + 5] += (ti || seen || inb);
                                extra2[8 * s + 6] += (ti || hitk[0]);
                            }
                            extra[4 * s + 0] += inb; extra[4 * s + 1] += seen; extra[4 * s + 2] += (inb || seen);
                        }
                        const uint32_t *v = tri + 3 * (size_t)t;
                        float dd = g3d_oracle_point_triangle_distance(gx, x + 3 * (size_t)v[0], x + 3 * (size_t)v[1], x + 3 * (size_t)v[2]);
                        if (dd < phi[q]) { phi[q] = dd; ct[q] = t; changed = 1; }
                    }
                    if (changed) { stamp[q] = (int8_t)(s + 1); ++o[4]; }
                    o[1] += active;
                    o[2] += evals > 0;
                    o[3] += evals;
                    round_active |= active;
                    level_active |= active;
                    ++cnt;
                    if (++in_round == 32) { ++o[5]; o[6] += round_active; in_round = 0; round_active = 0; }
                }
            if (in_round) { ++o[5]; o[6] += round_active; }
            o[7] += level_active;
        }
    }
    free(phi); free(ct); free(stamp);
}

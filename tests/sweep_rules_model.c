/* tests/sweep_rules_model.c -- TEST INFRASTRUCTURE. A CPU model of the shortcuts the CUDA sweep kernels take
 * (generate3d_b200/csrc/g3d_tdf.cu: change codes and re-examination thresholds, "met" neighbours, band-box flags,
 * speculative sweeps), written without any of the kernels' parallel machinery, so that tests/test_sweep_rules.py can
 * check on the CPU that every skipped evaluation really is a no-op: the model's phi and closest_tri must equal the
 * oracle's (oracle/g3d_oracle.c, the plain restatement of makelevelset3.cpp) bit for bit.
 *
 * Includes the oracle for the distance function and the band init. Built by the test with gcc -O2 -ffp-contract=off. */
#include "../oracle/g3d_oracle.c"

static void sdir(int s, int *d)              /* makelevelset3.cpp:163-172, pass 0 then pass 1 */
{
    int q8 = s & 7;
    d[0] = (q8 & 1) ? -1 : 1;
    d[1] = (q8 == 0 || q8 == 2 || q8 == 5 || q8 == 7) ? 1 : -1;
    d[2] = (q8 == 0 || q8 == 3 || q8 == 4 || q8 == 7) ? 1 : -1;
}

/* code of the last sweep before s that ran like s on the axes of `mask` (-1: none) */
static int threshold(int s, int mask)
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

typedef struct {
    const uint32_t *tri; const float *x; const float *origin; float dx;
    int ni, nj, nk;
    float *phi; int32_t *ct; int8_t *code; uint8_t *flags;
    int64_t evals;
} model;

/* One visit of node (i,j,k) in sweep s against the CURRENT state. apply = 0: only report whether it would change. */
static int visit(model *M, int s, int i, int j, int k, int apply)
{
    int d[3], dims[3] = { M->ni, M->nj, M->nk }, p[3] = { i, j, k };
    sdir(s, d);
    size_t q = (size_t)i + (size_t)M->ni * (j + (size_t)M->nj * k);
    int face = 0;
    for (int ax = 0; ax < 3; ++ax) if (p[ax] == 0 || p[ax] == dims[ax] - 1) face |= 1 << ax;
    int32_t t[7]; int met[7];
    for (int r = 0; r < 7; ++r) {
        int a = r + 1;
        int m[3] = { i - ((a & 1) ? d[0] : 0), j - ((a & 2) ? d[1] : 0), k - ((a & 4) ? d[2] : 0) };
        size_t qm = (size_t)m[0] + (size_t)M->ni * (m[1] + (size_t)M->nj * m[2]);
        t[r] = M->ct[qm];
        /* met: unchanged since this node last examined it, or still the band-init triangle whose band box holds this node */
        int need_flags = ((a & 1) ? (d[0] > 0 ? 1 : 2) : 0) | ((a & 2) ? (d[1] > 0 ? 4 : 8) : 0) | ((a & 4) ? (d[2] > 0 ? 16 : 32) : 0);
        met[r] = M->code[qm] <= threshold(s, a | face) || (M->code[qm] == 0 && t[r] >= 0 && (M->flags[qm] & need_flags) == need_flags);
    }
    float gx[3] = { i * M->dx + M->origin[0], j * M->dx + M->origin[1], k * M->dx + M->origin[2] };
    float phi = M->phi[q];
    int32_t ct = M->ct[q];
    const int32_t own = ct;
    int changed = 0;
    for (int r = 0; r < 7; ++r) {
        if (t[r] < 0 || met[r] || t[r] == own) continue;
        int drop = 0;
        for (int z = 0; z < r; ++z) if (t[z] == t[r]) drop = 1;             /* an earlier slot: evaluated first, or met before */
        for (int z = r + 1; z < 7; ++z) if (met[z] && t[z] == t[r]) drop = 1;   /* a later slot that was met before */
        if (drop) continue;
        const uint32_t *v = M->tri + 3 * (size_t)t[r];
        float dd = g3d_oracle_point_triangle_distance(gx, M->x + 3 * (size_t)v[0], M->x + 3 * (size_t)v[1], M->x + 3 * (size_t)v[2]);
        ++M->evals;
        if (dd < phi) { phi = dd; ct = t[r]; changed = 1; }
    }
    if (changed && apply) {
        /* (4-bit codes: the last sweep shares 15 with the one before; nothing reads codes after it) */
        M->phi[q] = phi; M->ct[q] = ct; M->code[q] = (int8_t)(s + 1 < 15 ? s + 1 : 15); M->flags[q] = 0;
    }
    return changed;
}

/* spec_from: sweeps >= spec_from run speculatively (99 = never). Returns the evaluations made in the sweeps. */
int64_t sweep_rules_model(const uint32_t *tri, int64_t ntri, const float *x, const float *origin, float dx, int ni, int nj,
                          int nk, int spec_from, float *phi_out, int32_t *ct_out)
{
    size_t n = (size_t)ni * nj * nk;
    model M = { tri, x, origin, dx, ni, nj, nk, phi_out, ct_out, (int8_t *)calloc(n, 1), (uint8_t *)calloc(n, 1), 0 };
    g3d_oracle_make_level_set3(tri, ntri, x, origin, dx, ni, nj, nk, phi_out, ct_out, NULL, 1, 1);   /* band init only */
    int dims[3] = { ni, nj, nk };
    for (size_t q = 0; q < n; ++q) {                 /* band-box flags of the nodes the band init gave a triangle */
        if (ct_out[q] < 0) continue;
        int p[3] = { (int)(q % ni), (int)((q / ni) % nj), (int)(q / ((size_t)ni * nj)) };
        const uint32_t *v = tri + 3 * (size_t)ct_out[q];
        for (int ax = 0; ax < 3; ++ax) {             /* cpp:131-137 */
            double f0 = ((double)x[3 * v[0] + ax] - origin[ax]) / dx, f1 = ((double)x[3 * v[1] + ax] - origin[ax]) / dx,
                   f2 = ((double)x[3 * v[2] + ax] - origin[ax]) / dx;
            int lo = clamp_i((int)min3_d(f0, f1, f2) - 1, 0, dims[ax] - 1), hi = clamp_i((int)max3_d(f0, f1, f2) + 2, 0, dims[ax] - 1);
            if (p[ax] + 1 <= hi) M.flags[q] |= 1 << (2 * ax);
            if (p[ax] - 1 >= lo) M.flags[q] |= 2 << (2 * ax);
        }
    }
    int na = ni - 1, nb = nj - 1, nc = nk - 1, nlev = na + nb + nc - 2;
    uint8_t *mark = (uint8_t *)calloc(n, 1);
    for (int s = 0; s < 16 && nlev > 0; ++s) {
        int d[3];
        sdir(s, d);
        const int spec = s >= spec_from;
        if (spec) {
            /* every node at once against the sweep's initial state: which would change? (nothing is stored) */
            memset(mark, 0, n);
            for (int c = 0; c < nc; ++c) for (int b = 0; b < nb; ++b) for (int a = 0; a < na; ++a) {
                int i = d[0] > 0 ? 1 + a : ni - 2 - a, j = d[1] > 0 ? 1 + b : nj - 2 - b, k = d[2] > 0 ? 1 + c : nk - 2 - c;
                if (visit(&M, s, i, j, k, 0)) mark[(size_t)i + (size_t)ni * (j + (size_t)nj * k)] = 1;
            }
        }
        /* level by level (a + b + c), inside a level back to front: the order inside a level must not matter */
        for (int L = 0; L < nlev; ++L)
            for (int c = nc - 1; c >= 0; --c) for (int b = nb - 1; b >= 0; --b) {
                int a = L - b - c;
                if (a < 0 || a >= na) continue;
                int i = d[0] > 0 ? 1 + a : ni - 2 - a, j = d[1] > 0 ? 1 + b : nj - 2 - b, k = d[2] > 0 ? 1 + c : nk - 2 - c;
                size_t q = (size_t)i + (size_t)ni * (j + (size_t)nj * k);
                if (spec && !mark[q]) continue;      /* would not change, and nothing upwind of it has changed */
                if (visit(&M, s, i, j, k, 1) && spec) {
                    for (int a2 = 1; a2 < 8; ++a2) {  /* the 7 nodes downwind of a change have a new candidate */
                        int ii = i + ((a2 & 1) ? d[0] : 0), jj = j + ((a2 & 2) ? d[1] : 0), kk = k + ((a2 & 4) ? d[2] : 0);
                        if (ii < 0 || jj < 0 || kk < 0 || ii >= ni || jj >= nj || kk >= nk) continue;
                        mark[(size_t)ii + (size_t)ni * (jj + (size_t)nj * kk)] = 1;
                    }
                }
            }
    }
    free(M.code); free(M.flags); free(mark);
    return M.evals;
}

// g3d_tdf_exact.cuh -- "exact" mode: nearest triangle per voxel through a per-mesh cell grid.
// Included by g3d_tdf.cu (shares tdf_prep / tdf_finalize and the key format).
//
// The reference's far field is only approximate (makelevelset3.h:7-11: beyond the 1-cell band a
// distance "might not be to the closest triangle"). This mode instead takes, for every voxel, the
// true minimum of the reference's own float distance function over ALL triangles of the mesh --
// bit-equal to the reference wherever the reference is exact (|d| <= 1 voxel, hence identical
// occupancy), never larger elsewhere -- for every voxel whose distance can matter:
//   * outside voxels (even x-ray parity): exact when d <= trunc (anything farther is clamped to
//     trunc by load_df, dataset_loader.py:137-139) and reported as the reference's initial "far"
//     value (ni+nj+nk)*dx otherwise;
//   * inside voxels (odd parity): always exact (the negative side is never clamped).
//
// Spatial index (built per call, in voxel-index space, no host involvement):
//   tdf_prep      per triangle: bounding box in grid coordinates, rounded OUTWARD to half precision; home cell =
//                 the 2x2x2-node cell holding the box centre. A triangle whose box stays within its home cell
//                 dilated by kSmallMax voxels is "small" and is counted into that cell; anything bigger goes to
//                 the mesh's "large" list. The largest overhang of a small triangle is kept per mesh (dil).
//   exact_scan    per mesh: exclusive scan of the cell counts -> CSR starts (the large list is pseudo-cell ncell)
//   exact_fill    per triangle: entry {box (6 halves), mesh-local triangle index} to its CSR slot (16 B)
//   exact_cellbox per cell: union of its entries' boxes (16 B: what a group tests before touching the list)
//   tdf_exact3    one WARP per cell = per 2x2x2 group of voxels (lane = voxel x 4 triangle slots). The warp first
//                 scans the large list, then its own cell, then the surrounding cells shell by shell (Chebyshev
//                 rings of cells), inside a ring nearest cell box first, and stops at the first ring that cannot
//                 reach below the group's running bound. Because the nearest cells come first the bound is tight
//                 after the first few triangles, so a voxel meets a few dozen candidates instead of thousands.
//   Per list, 32 entries at a time: entry box vs the group's box with the running bound -> survivors (ballot) ->
//   four survivors at a time are rated against the 8 voxels (stage 1, cheap FMA squared distance with an error
//   interval, no division or square root) -> the few pairs whose interval reaches below the voxel's running upper
//   bound are queued and re-evaluated 32 at a time with the bit-faithful function (stage 2), which alone decides
//   the result; a warp-shuffle min-reduce shares the running bounds between the four slots of a voxel and the
//   eight voxels of the group.
#pragma once
// (needs <cuda_fp16.h>, included by g3d_tdf.cu at global scope: this file is included inside namespace g3d)

namespace exact {

constexpr int kCell = 2;                  // nodes per cell and axis (one warp-group of 2x2x2 voxels per cell)
constexpr float kSmallMax = 2.0f;         // overhang (voxels) beyond the home cell up to which a triangle is "small"
constexpr uint32_t kNoCell = 0xffffffffu;

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---------------------------------------------------------------------------------------------
// Two-stage evaluation. Stage 1 rates a (voxel, triangle) pair with a cheap squared distance that is free to use
// FMA (55 arithmetic instructions, no division, no square root: the three clamped edge parameters use precomputed
// reciprocals and the saturate modifier). Only pairs whose cheap distance is within a generous margin of the
// voxel's running cheap minimum can be the minimum of the reference's float function; those few are queued and
// re-evaluated 32 at a time with the bit-faithful function (stage 2), which alone decides the result.
//
// Why this is still exact: stage 1 works on intervals. For every non-flagged triangle the cheap squared
// distance a2 comes with an error bound e2 = 2e-6 * (|x0-x3|^2 + the three squared edge lengths): every
// intermediate of the rating is bounded by that sum and fewer than twenty roundings of 2^-24 relative size are
// involved, and the closest point is formed explicitly (x3 + w23 x13 + w31 x23) so that the error of the
// barycentric solve enters only to second order (slivers with sin^2 < 0.001 are flagged instead). So
// a2 - e2 <= d^2 <= a2 + e2 for the true distance d, and the reference's float function f agrees with d to
// about 1e-7 relative (measured; its own closest-point construction has the same second-order property).
// A voxel keeps ub2 = min_t (a2 + e2), an upper bound of min_t d^2, and a triangle goes to stage 2 iff
// a2 - e2 <= 1.002 * ub2 + kappa. The minimiser t0 of f always does: d(t0) <= f(t0)(1+1e-6) <= f(t)(1+1e-6)
// <= d(t)(1+2e-6) for every t, hence a2(t0) - e2(t0) <= d(t0)^2 <= 1.00001 * ub2. kappa (1e-10 * S^2, S the
// extent of the coordinates) covers distances so small that f's absolute rounding dominates. Triangles for which
// no such bound holds -- slivers, zero-length edges, non-finite vertices -- are flagged at record-build time and
// always go to stage 2. The box culls in front of stage 1 compare a LOWER bound of d^2 (distance between the
// outward-rounded boxes) with 1.001 * the largest lim2 of the group, so they never drop a possible minimiser.
struct FastRec {                                     // 5 x float4 = 80 B
    float4 f0;   // x3.xyz, invdet (NaN = always evaluate faithfully)
    float4 f1;   // x13.xyz, m13
    float4 f2;   // x23.xyz, m23
    float4 f3;   // e12.xyz, m12
    float4 f4;   // d, 1/m13, 1/m23, 1/m12
};

__device__ __forceinline__ void store_fastrec(float4* p, V3 x1, V3 x2, V3 x3)
{
    const float qn = __int_as_float(0x7fc00000);
    V3 x13 = { x1.x - x3.x, x1.y - x3.y, x1.z - x3.z }, x23 = { x2.x - x3.x, x2.y - x3.y, x2.z - x3.z };
    V3 e12 = { x2.x - x1.x, x2.y - x1.y, x2.z - x1.z };
    float m13 = x13.x * x13.x + x13.y * x13.y + x13.z * x13.z;
    float m23 = x23.x * x23.x + x23.y * x23.y + x23.z * x23.z;
    float m12 = e12.x * e12.x + e12.y * e12.y + e12.z * e12.z;
    float d = x13.x * x23.x + x13.y * x23.y + x13.z * x23.z;
    float det = m13 * m23 - d * d;
    bool ok = (m13 > 0.f) && (m23 > 0.f) && (m12 > 0.f) && (det > 1e-3f * m13 * m23) && (fabsf(m13 + m23 + m12) < 1e30f) &&
              (fabsf(x3.x) + fabsf(x3.y) + fabsf(x3.z) < 1e15f);
    p[0] = make_float4(x3.x, x3.y, x3.z, ok ? 1.f / det : qn);
    p[1] = make_float4(x13.x, x13.y, x13.z, m13);
    p[2] = make_float4(x23.x, x23.y, x23.z, m23);
    p[3] = make_float4(e12.x, e12.y, e12.z, m12);
    p[4] = make_float4(d, 1.f / m13, 1.f / m23, 1.f / m12);
}

// cheap squared point-triangle distance a2 and its error bound e2; `flagged` is set for records that must go
// to stage 2 unconditionally
__device__ __forceinline__ float fast_dist2(V3 x0, const float4* __restrict__ p, bool& flagged, float& e2)
{
    const float4 f0 = __ldg(p), f1 = __ldg(p + 1), f2 = __ldg(p + 2), f3 = __ldg(p + 3), f4 = __ldg(p + 4);
    flagged = !(f0.w == f0.w);
    const float vx = x0.x - f0.x, vy = x0.y - f0.y, vz = x0.z - f0.z;
    const float a = fmaf(f1.z, vz, fmaf(f1.y, vy, f1.x * vx));
    const float b = fmaf(f2.z, vz, fmaf(f2.y, vy, f2.x * vx));
    const float vv = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
    const float w23 = f0.w * fmaf(-f4.x, b, f2.w * a);
    const float w31 = f0.w * fmaf(-f4.x, a, f1.w * b);
    const float w12 = 1.f - w23 - w31;
    // x0 minus its projection onto the plane, formed explicitly: errors of w23 / w31 move the foot point inside
    // the plane only, which changes the length of this vector to second order
    const float px = fmaf(-w31, f2.x, fmaf(-w23, f1.x, vx));
    const float py = fmaf(-w31, f2.y, fmaf(-w23, f1.y, vy));
    const float pz = fmaf(-w31, f2.z, fmaf(-w23, f1.z, vz));
    const float plane = fmaf(pz, pz, fmaf(py, py, px * px));
    const float s1 = __saturatef(a * f4.y), s2 = __saturatef(b * f4.z);
    const float e1 = fmaf(s1, fmaf(s1, f1.w, -2.f * a), vv);
    const float e2b = fmaf(s2, fmaf(s2, f2.w, -2.f * b), vv);
    const float ux = vx - f1.x, uy = vy - f1.y, uz = vz - f1.z;       // x0 - x1
    const float c = fmaf(f3.z, uz, fmaf(f3.y, uy, f3.x * ux));
    const float uu = fmaf(uz, uz, fmaf(uy, uy, ux * ux));
    const float s3 = __saturatef(c * f4.w);
    const float e3 = fmaf(s3, fmaf(s3, f3.w, -2.f * c), uu);
    const float edge = fminf(e1, fminf(e2b, e3));
    const bool inside = (w23 >= 0.f) && (w31 >= 0.f) && (w12 >= 0.f);
    e2 = 2e-6f * (vv + f1.w + f2.w + f3.w);
    return fmaxf(inside ? plane : edge, 0.f);
}

// ---------------------------------------------------------------------------------------------
// boxes in voxel-index space, six halves in three words: (lo.x, lo.y) (lo.z, hi.x) (hi.y, hi.z)
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b)
{
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t w)
{
    return __half22float2(__halves2half2(__ushort_as_half((unsigned short)(w & 0xffffu)), __ushort_as_half((unsigned short)(w >> 16))));
}
// outward rounding: the half box contains the double box
__device__ __forceinline__ uint3 pack_box(const double* mn, const double* mx)
{
    __half l[3], h[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        l[a] = __float2half_rd(__double2float_rd(mn[a]));
        h[a] = __float2half_ru(__double2float_ru(mx[a]));
    }
    return make_uint3(pack_h2(l[0], l[1]), pack_h2(l[2], h[0]), pack_h2(h[1], h[2]));
}
// squared distance (voxel units) between a packed box and the box [glo, ghi]; 0 when they overlap
__device__ __forceinline__ float box_gap2(uint32_t w0, uint32_t w1, uint32_t w2, float3 glo, float3 ghi)
{
    const float2 a = unpack_h2(w0), b = unpack_h2(w1), c = unpack_h2(w2);   // lo.x lo.y | lo.z hi.x | hi.y hi.z
    const float gx = fmaxf(0.f, fmaxf(a.x - ghi.x, glo.x - b.y));
    const float gy = fmaxf(0.f, fmaxf(a.y - ghi.y, glo.y - c.x));
    const float gz = fmaxf(0.f, fmaxf(b.x - ghi.z, glo.z - c.y));
    return fmaf(gx, gx, fmaf(gy, gy, gz * gz));
}

struct Ws {
    float4* tri; float4* fast; uint4* tmp; uint4* ent; u64* key; uint32_t* parity;
    uint32_t* cell_start;     // [n_mesh][ncell + 2]: CSR starts relative to the mesh's first entry; [ncell] = large list, [ncell+1] = end
    uint32_t* cursor;         // [n_mesh][ncell + 1]: counts (prep), then fill cursors
    uint4* cellbox;           // [n_mesh][ncell]: union box of the cell's entries (3 words) + spare
    uint32_t* dil;            // [n_mesh]: float bits of the largest overhang of a small triangle beyond its home cell
    uint32_t* occ;            // [n_mesh][ceil(ncell / 32)]: bit = the cell holds at least one entry
    size_t bytes;
};

__host__ __device__ inline int cells_of(int n) { return (n + kCell - 1) / kCell; }
__host__ __device__ inline int64_t ncell_of(const g3d_tdf_params& p) { return (int64_t)cells_of(p.ni) * cells_of(p.nj) * cells_of(p.nk); }

inline Ws carve(void* base, int32_t n_mesh, int64_t total_tris, const g3d_tdf_params& p)
{
    Ws w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (char*)base + o; };
    const int64_t nvox = nvox_of(p), pw = (nvox + 31) / 32, ncell = ncell_of(p);
    w.tri = (float4*)take((size_t)total_tris * 64);
    w.fast = (float4*)take((size_t)total_tris * 80);
    w.tmp = (uint4*)take((size_t)total_tris * 16);
    w.ent = (uint4*)take((size_t)total_tris * 16);
    w.key = (u64*)take((size_t)n_mesh * nvox * 8);
    w.parity = (uint32_t*)take((size_t)n_mesh * pw * 4);
    w.cell_start = (uint32_t*)take((size_t)n_mesh * (ncell + 2) * 4);
    w.cursor = (uint32_t*)take((size_t)n_mesh * (ncell + 1) * 4);
    w.cellbox = (uint4*)take((size_t)n_mesh * ncell * 16);
    w.dil = (uint32_t*)take((size_t)n_mesh * 4);
    w.occ = (uint32_t*)take((size_t)n_mesh * ((ncell + 31) / 32) * 4);
    w.bytes = off;
    return w;
}

// the per-triangle part of the index, called by tdf_prep for a valid, finite triangle; f* are its corners in
// grid coordinates (double, makelevelset3.cpp:131-133)
__device__ __forceinline__ void index_triangle(const double* fp, const double* fq, const double* fr, int m, int64_t t,
                                               const g3d_tdf_params& P, uint4* tmp, uint32_t* cursor, uint32_t* dil)
{
    const int nc[3] = { cells_of(P.ni), cells_of(P.nj), cells_of(P.nk) };
    double mn[3], mx[3];
    int c[3];
    double over = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mn[a] = fmin(fp[a], fmin(fq[a], fr[a]));
        mx[a] = fmax(fp[a], fmax(fq[a], fr[a]));
        double mid = 0.5 * (mn[a] + mx[a]) / (double)kCell;
        mid = mid < 0.0 ? 0.0 : (mid > (double)(nc[a] - 1) ? (double)(nc[a] - 1) : mid);
        c[a] = (int)mid;                                         // floor, mid >= 0
        const double lo = (double)(kCell * c[a]), hi = lo + (double)kCell;
        over = fmax(over, fmax(lo - mn[a], mx[a] - hi));
    }
    const int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
    const bool small = over <= (double)kSmallMax;
    const int64_t cell = small ? (c[0] + (int64_t)nc[0] * (c[1] + (int64_t)nc[1] * c[2])) : ncell;
    if (small) {                                                 // non-negative floats order like uints
        const uint32_t ob = __float_as_uint(__double2float_ru(over < 0.0 ? 0.0 : over));
        if (ob > *(volatile uint32_t*)(dil + m)) atomicMax(dil + m, ob);      // all triangles of a mesh share this word: look first
    }
    atomicAdd(cursor + (int64_t)m * (ncell + 1) + cell, 1u);
    const uint3 b = pack_box(mn, mx);
    tmp[t] = make_uint4(b.x, b.y, b.z, (uint32_t)cell);
}

// one CTA per mesh: counts -> exclusive starts, tile by tile with coalesced accesses (a 128^3 grid has 262,145 counts);
// the counters become fill cursors (zero)
__global__ void __launch_bounds__(1024)
exact_scan_kernel(uint32_t* __restrict__ cursor, uint32_t* __restrict__ cell_start, int64_t ncell)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int mesh = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t* cnt = cursor + (int64_t)mesh * (ncell + 1);
    uint32_t* st = cell_start + (int64_t)mesh * (ncell + 2);
    const int64_t n = ncell + 1;
    if (tid == 0) s_carry = 0u;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t q = base + tid;
        const uint32_t v = q < n ? cnt[q] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_warp[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += up;
            }
            s_warp[lane] = wi - w;                               // exclusive offset of each warp inside the tile
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + s_warp[wid] + incl - v;
        if (q < n) { st[q] = excl; cnt[q] = 0u; }
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;                     // inclusive end of the tile
        __syncthreads();
    }
    if (tid == 0) st[n] = s_carry;
}

__global__ void __launch_bounds__(256)
exact_fill_kernel(const uint4* __restrict__ tmp, const int64_t* __restrict__ tri_off, int64_t total_tris, int n_mesh,
                  int64_t ncell, const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cursor, uint4* __restrict__ ent)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_tris) return;
    const uint4 e = tmp[t];
    if (e.w == kNoCell) return;
    int lo = 0, hi = n_mesh;                                     // tri_off[lo] <= t < tri_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tri_off[mid] <= t) lo = mid; else hi = mid;
    }
    const int64_t tb = tri_off[lo];
    const uint32_t pos = cell_start[(int64_t)lo * (ncell + 2) + e.w] + atomicAdd(cursor + (int64_t)lo * (ncell + 1) + e.w, 1u);
    ent[tb + pos] = make_uint4(e.x, e.y, e.z, (uint32_t)(t - tb));
}

// one thread per cell: union of the cell's entry boxes (min / max of halves is exact)
__global__ void __launch_bounds__(256)
exact_cellbox_kernel(const uint4* __restrict__ ent, const int64_t* __restrict__ tri_off, const uint32_t* __restrict__ cell_start,
                     int64_t ncell, int64_t total_cells, uint4* __restrict__ cellbox, uint32_t* __restrict__ occ)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total_cells) return;
    const int mesh = (int)(q / ncell);
    const int64_t c = q - (int64_t)mesh * ncell;
    const uint32_t* st = cell_start + (int64_t)mesh * (ncell + 2);
    const uint32_t s = st[c], e = st[c + 1];
    const float inf = __int_as_float(0x7f800000);
    float lx = inf, ly = inf, lz = inf, hx = -inf, hy = -inf, hz = -inf;
    const uint4* E = ent + tri_off[mesh];
    for (uint32_t p = s; p < e; ++p) {
        const uint4 en = __ldg(E + p);
        const float2 a = unpack_h2(en.x), b = unpack_h2(en.y), d = unpack_h2(en.z);
        lx = fminf(lx, a.x); ly = fminf(ly, a.y); lz = fminf(lz, b.x);
        hx = fmaxf(hx, b.y); hy = fmaxf(hy, d.x); hz = fmaxf(hz, d.y);
    }
    if (e > s) atomicOr(occ + (int64_t)mesh * ((ncell + 31) / 32) + (c >> 5), 1u << (c & 31));
    cellbox[q] = make_uint4(pack_h2(__float2half_rd(lx), __float2half_rd(ly)), pack_h2(__float2half_rd(lz), __float2half_ru(hx)),
                            pack_h2(__float2half_ru(hy), __float2half_ru(hz)), e - s);
}

// offset (relative to the block's first cell) of the q-th cell of Chebyshev ring k around a block of CX x CY x CZ cells.
// Ring 0 is the block itself; ring k >= 1 is the shell of the block grown by k cells: the two z faces, then the y faces
// without their z edges, then the x faces without their y and z edges.
template <int CX, int CY, int CZ>
__device__ __forceinline__ int ring_size(int k)
{
    const int w = CX + 2 * k, h = CY + 2 * k, d = CZ + 2 * k;
    return k == 0 ? CX * CY * CZ : w * h * d - (w - 2) * (h - 2) * (d - 2);
}
template <int CX, int CY, int CZ>
__device__ __forceinline__ void ring_offset(int k, int q, int& dx, int& dy, int& dz)
{
    if (k == 0) { dx = q % CX; dy = (q / CX) % CY; dz = q / (CX * CY); return; }
    const int w = CX + 2 * k, h = CY + 2 * k, d = CZ + 2 * k, A = w * h, B = w * (d - 2), C = (h - 2) * (d - 2);
    if (q < 2 * A) {
        const int s = q >= A, r = q - s * A;
        dz = s ? CZ - 1 + k : -k; dy = r / w - k; dx = r - (r / w) * w - k;
    } else if (q < 2 * A + 2 * B) {
        q -= 2 * A;
        const int s = q >= B, r = q - s * B;
        dy = s ? CY - 1 + k : -k; dz = r / w - (k - 1); dx = r - (r / w) * w - k;
    } else {
        q -= 2 * A + 2 * B;
        const int s = q >= C, r = q - s * C;
        dx = s ? CX - 1 + k : -k; dz = r / (h - 2) - (k - 1); dy = r - (r / (h - 2)) * (h - 2) - (k - 1);
    }
}

constexpr int kExactWarps = 4;            // warps per CTA

// A warp owns a GROUP of CX x CY x CZ cells = NV = 8 CX CY CZ voxels (NV <= 32; lane = voxel wherever voxels are
// enumerated). Bigger groups share the walk of the index and the entry tests between more voxels at the price of a
// looser group bound (measured in tdf_solve's launcher).
template <int MINB, int CX, int CY, int CZ>
__global__ void __launch_bounds__(kExactWarps * 32, MINB)
tdf_exact3_kernel(const float4* __restrict__ tri, const float4* __restrict__ fast, const uint4* __restrict__ ent,
                  const uint32_t* __restrict__ cell_start, const uint4* __restrict__ cellbox, const uint32_t* __restrict__ occ,
                  const uint32_t* __restrict__ dil, const int64_t* __restrict__ tri_off, const uint32_t* __restrict__ parity,
                  u64* __restrict__ key, g3d_tdf_params P, float R, float far, float kappa, int64_t n_groups, u64* counter)
{
    constexpr int GX = kCell * CX, GY = kCell * CY, GZ = kCell * CZ, NV = GX * GY * GZ;
    static_assert(NV <= 32, "a group is at most one voxel per lane");
    constexpr int R1 = (CX + 2) * (CY + 2) * (CZ + 2) - CX * CY * CZ, R2 = (CX + 4) * (CY + 4) * (CZ + 4) - (CX + 2) * (CY + 2) * (CZ + 2);
    constexpr int kRingTable = CX * CY * CZ + R1 + R2;     // rings 0..2 from a shared-memory table
    static_assert((CY + 4) * (CZ + 4) <= 32 && CX + 4 <= 32, "the occupancy pre-test gives one (y, z) row of cells to a lane");
    // Two warp-private queues turn two uneven filters into dense work:
    //   pairs  (voxel, triangle) pairs that passed the per-voxel box test, waiting for stage 1 (32 at a time, one pair per lane)
    //   queue  pairs whose stage-1 interval reaches below the voxel's bound, waiting for stage 2 (32 at a time)
    // The group's work is ONE loop -- pick the next list, take 32 entries, test each entry's box against the group's
    // voxels, run whichever stage has a full batch -- so that each stage exists once in the instruction stream (inlined at
    // every call site the kernel was twice the 32 KB instruction cache and mostly waited for instructions).
    __shared__ uint32_t s_pairs[kExactWarps][64];
    __shared__ uint32_t s_queue[kExactWarps][64];
    __shared__ uint32_t s_ub[kExactWarps][32];       // per voxel of the group: float bits of ub2, the smallest a2 + e2 seen
    __shared__ __align__(16) float s_vb[kExactWarps][32];   // per voxel: the cull bound derived from ub2 (voxel units^2; -1 outside the grid)
    __shared__ u64 s_best[kExactWarps][32];          // per voxel: float bits of the best faithful distance << 32 | its triangle
    __shared__ int s_ring[kRingTable];               // packed offsets (dx+8) | (dy+8)<<4 | (dz+8)<<8 of rings 0, 1 and 2
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (int q = threadIdx.x; q < kRingTable; q += kExactWarps * 32) {
        int dx, dy, dz;
        if (q >= CX * CY * CZ + R1) ring_offset<CX, CY, CZ>(2, q - CX * CY * CZ - R1, dx, dy, dz);
        else if (q >= CX * CY * CZ) ring_offset<CX, CY, CZ>(1, q - CX * CY * CZ, dx, dy, dz);
        else ring_offset<CX, CY, CZ>(0, q, dx, dy, dz);
        s_ring[q] = (dx + 8) | ((dy + 8) << 4) | ((dz + 8) << 8);
    }
    __syncthreads();
    uint32_t* pairs = s_pairs[warp];
    uint32_t* qu = s_queue[warp];
    volatile uint32_t* ubv = s_ub[warp];
    volatile float* vbs = s_vb[warp];
    const int ni = P.ni, nj = P.nj, nk = P.nk;
    const int ncx = cells_of(ni), ncy = cells_of(nj), ncz = cells_of(nk);
    const int ngx = (ncx + CX - 1) / CX, ngy = (ncy + CY - 1) / CY, ngz = (ncz + CZ - 1) / CZ;
    const int64_t ncell = (int64_t)ncx * ncy * ncz, ngroup = (int64_t)ngx * ngy * ngz, nvox = nvox_of(P), occ_words = (ncell + 31) / 32;
    const float inv_dx = 1.0f / P.dx;
    const float to_vox2 = inv_dx * inv_dx * 1.0001f * 1.001f;    // world distance^2 -> voxel units, rounded up, with the cull margin
    const float INF = __int_as_float(0x7f800000);
    u64 n_pairs = 0, n_tests = 0, n_faithful = 0;
    enum { ST_LARGE, ST_BATCH, ST_REST, ST_RINGEND, ST_FINAL };

    const int64_t wstride = (int64_t)gridDim.x * kExactWarps;
    for (int64_t g = (int64_t)blockIdx.x * kExactWarps + warp; g < n_groups; g += wstride) {
        // (32-bit divisions whenever the numbers allow: a 64-bit division is ~100 instructions, and there were five per group)
        int mesh, cx, cy, cz;
        if (n_groups < 0x7fffffffll) {
            const uint32_t g32 = (uint32_t)g, ng32 = (uint32_t)ngroup;
            mesh = (int)(g32 / ng32);
            const uint32_t cg = g32 - (uint32_t)mesh * ng32, row = cg / (uint32_t)ngx;
            cx = (int)(cg - row * (uint32_t)ngx) * CX;
            cz = (int)(row / (uint32_t)ngy);
            cy = (int)(row - (uint32_t)cz * (uint32_t)ngy) * CY;
            cz *= CZ;
        } else {
            mesh = (int)(g / ngroup);
            const int64_t cg = g - (int64_t)mesh * ngroup;
            cx = (int)(cg % ngx) * CX; cy = (int)((cg / ngx) % ngy) * CY; cz = (int)(cg / ((int64_t)ngx * ngy)) * CZ;
        }
        // (cx, cy, cz): first cell of the group's block of cells; the group's first voxel is kCell times that
        // lanes 0..NV-1 own the group's voxels (lane = voxel): validity, sign, start value, result
        const int i = kCell * cx + lane % GX, j = kCell * cy + (lane / GX) % GY, k = kCell * cz + lane / (GX * GY);
        const bool valid = lane < NV && i < ni && j < nj && k < nk;
        bool inside = false;
        if (valid) {                                    // running parity along i (makelevelset3.cpp:174-182)
            const uint32_t* par = parity + (int64_t)mesh * ((nvox + 31) / 32);
            const int64_t r0 = (int64_t)ni * (j + (int64_t)nj * k), r1 = r0 + i;
            int cnt = 0;
            for (int64_t w = r0 >> 5; w <= (r1 >> 5); ++w) {
                uint32_t m = par[w];
                if (w == (r0 >> 5)) m &= 0xffffffffu << (r0 & 31);
                if (w == (r1 >> 5)) m &= 0xffffffffu >> (31 - (r1 & 31));
                cnt += __popc(m);
            }
            inside = cnt & 1;
        }
        // like the reference's phi, the running minimum starts at `far` (makelevelset3.cpp:123): a triangle farther
        // away than that never registers. Outside voxels may stop at the truncation radius instead.
        const float start = inside ? far : fminf(R, far);
        if (!__any_sync(0xffffffffu, valid)) continue;  // cannot happen for groups inside the grid
        // stage 1 passes a pair to stage 2 iff a2 - e2 <= lim2 = 1.002 * ub2 + kappa, ub2 = smallest a2 + e2 seen (start^2 at first).
        // stage 2 keeps (distance, triangle) as one 64-bit key per voxel: atomicMin is "smaller distance, then lower index";
        // the initial key (start, 0) can only be beaten by a distance strictly below start, like the reference's d < phi.
        const u64 key0 = (u64)__float_as_uint(start) << 32;
        __syncwarp();
        ubv[lane] = __float_as_uint(start * start);
        s_best[warp][lane] = key0;
        // the per-voxel cull bounds (shared memory, one per lane's voxel) and their maximum over the group
        float gb2;
        auto load_bounds = [&]() {
            const float b = valid ? fmaf(__uint_as_float(ubv[lane]), 1.002f, kappa) * to_vox2 : -1.f;
            vbs[lane] = b;
            gb2 = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(b, 0.f))));   // non-negative floats order like uints
            __syncwarp();
        };
        __syncwarp();
        load_bounds();
        const float x0 = (float)(kCell * cx), y0 = (float)(kCell * cy), z0 = (float)(kCell * cz);
        const float3 glo = make_float3(x0, y0, z0);
        const float3 ghi = make_float3((float)min(kCell * cx + GX - 1, ni - 1), (float)min(kCell * cy + GY - 1, nj - 1),
                                       (float)min(kCell * cz + GZ - 1, nk - 1));

        const int64_t t0 = tri_off[mesh];
        const float4* Tm = tri + 4 * t0;
        const float4* Fm = fast + 5 * t0;
        const uint4* Em = ent + t0;
        const uint32_t* CS = cell_start + (int64_t)mesh * (ncell + 2);
        const uint4* CB = cellbox + (int64_t)mesh * ncell;
        // a small triangle overhangs its home cell by at most D voxels (+ the outward rounding of the half-precision
        // boxes: one ulp at the grid's far end, 11 significant bits); ring k's cells are at least 2k - 2 voxels away
        const float D = __uint_as_float(__ldg(dil + mesh)) * 1.01f + 0.05f + (float)max(ni, max(nj, nk)) * (1.01f / 1024.f);
        const int kmax = max(max(cx, ncx - cx - CX), max(max(cy, ncy - cy - CY), max(cz, ncz - cz - CZ)));

        // anything at all nearby? the large list's length and the occupancy bits of the cells of rings 0..2 (one (y, z) row of
        // cells per lane) decide whether those rings need a visit
        uint32_t n_large = 0;
        bool near = false;
        {
            bool hit = false;
            if (lane < (CY + 4) * (CZ + 4)) {
                const int y = cy + lane % (CY + 4) - 2, z = cz + lane / (CY + 4) - 2;
                if (y >= 0 && z >= 0 && y < ncy && z < ncz) {
                    const uint32_t* OC = occ + (int64_t)mesh * occ_words;
                    const int xs = max(cx - 2, 0), nb = min(cx + CX + 1, ncx - 1) - xs + 1;
                    const int64_t b0 = xs + (int64_t)ncx * (y + (int64_t)ncy * z);
                    const int sh = (int)(b0 & 31);
                    uint32_t w = __ldg(OC + (b0 >> 5)) >> sh;
                    if (sh + nb > 32) w |= __ldg(OC + (b0 >> 5) + 1) << (32 - sh);
                    hit = (w & ((1u << nb) - 1u)) != 0u;
                }
            }
            near = __any_sync(0xffffffffu, hit);
            if (lane == 0) n_large = __ldg(CS + ncell + 1) - __ldg(CS + ncell);
            n_large = __shfl_sync(0xffffffffu, n_large, 0);
        }

        int pqn = 0, qn = 0;                             // fills of pairs / queue, warp-uniform
        // the list being walked: the concatenation of one list per lane (start st, length cnt, exclusive offset excl)
        uint32_t st = 0, cnt = 0, excl = 0, total = 0, pos = 0;
        int single = -1;
        uint32_t rest_cnt = 0;                           // the batch's other cells, admitted after the nearest one
        float rest_g2 = INF;
        int state = n_large ? ST_LARGE : ST_BATCH, kr = near ? 0 : 3, q0 = 0;
        if (!near) {
            const float lb3 = 4.f - D;
            if (kr > kmax || (lb3 > 0.f && lb3 * lb3 > gb2)) state = n_large ? ST_LARGE : ST_FINAL;
        }
        bool drained = state == ST_FINAL;                // nothing was queued yet: no drain needed to finish
        for (;;) {
            bool drain = false, final = false;
            uint32_t vm = 0, tid = 0;                    // this lane's entry: voxels (bit mask) whose box test it passed, triangle
            if (pos >= total) {
                // ---- pick the next list (warp-uniform control flow)
                bool fresh = false;
                if (state == ST_LARGE) {                 // triangles too big for the grid: every group looks at all of them
                    cnt = 0;
                    if (lane == 0) { st = __ldg(CS + ncell); cnt = n_large; }
                    state = ST_BATCH;
                    if (!near) {
                        const float lb3 = 4.f - D;
                        if (kr > kmax || (lb3 > 0.f && lb3 * lb3 > gb2)) state = ST_FINAL;
                    }
                    fresh = true;
                } else if (state == ST_BATCH) {
                    const int n = ring_size<CX, CY, CZ>(kr);
                    if (q0 >= n) {
                        state = ST_RINGEND;
                    } else {
                        // up to 32 cells of ring kr (lane = cell): boxes and list starts of the whole batch in one round trip
                        const int q = q0 + lane;
                        int dx = 0, dy = 0, dz = 0;
                        if (q < n) {
                            if (kr <= 2) {
                                const int o = s_ring[(kr == 2 ? CX * CY * CZ + R1 : kr == 1 ? CX * CY * CZ : 0) + q];
                                dx = (o & 15) - 8; dy = ((o >> 4) & 15) - 8; dz = ((o >> 8) & 15) - 8;
                            } else {
                                ring_offset<CX, CY, CZ>(kr, q, dx, dy, dz);
                            }
                        }
                        const int ccx = cx + dx, ccy = cy + dy, ccz = cz + dz;
                        uint32_t keyv = 0xffffffffu;
                        cnt = 0; rest_cnt = 0; rest_g2 = INF;
                        if (q < n && ccx >= 0 && ccy >= 0 && ccz >= 0 && ccx < ncx && ccy < ncy && ccz < ncz) {
                            const int64_t cid = ccx + (int64_t)ncx * (ccy + (int64_t)ncy * ccz);
                            const uint4 cb = __ldg(CB + cid);
                            st = __ldg(CS + cid);
                            ++n_tests;
                            if (cb.w) {
                                rest_g2 = box_gap2(cb.x, cb.y, cb.z, glo, ghi);
                                if (rest_g2 <= gb2) { keyv = (__float_as_uint(rest_g2) & ~31u) | (uint32_t)lane; rest_cnt = cb.w; }
                            }
                        }
                        const uint32_t kmin = __reduce_min_sync(0xffffffffu, keyv);
                        if (kmin == 0xffffffffu) {
                            q0 += 32;                    // nothing of this batch is within reach
                        } else {                         // the nearest cell first, on its own: it tightens the bound
                            if (lane == (int)(kmin & 31u)) { cnt = rest_cnt; rest_cnt = 0; }
                            state = __any_sync(0xffffffffu, rest_cnt != 0) ? ST_REST : ST_BATCH;
                            if (state == ST_BATCH) q0 += 32;
                            fresh = true;
                        }
                    }
                } else if (!drained) {                   // ST_REST / ST_RINGEND / ST_FINAL: everything waiting goes through first
                    drain = true;
                    final = state == ST_FINAL;
                    drained = true;
                } else {
                    drained = false;
                    if (state == ST_REST) {              // the other cells of the batch, against the tightened bound
                        cnt = rest_g2 <= gb2 ? rest_cnt : 0u;
                        state = ST_BATCH;
                        q0 += 32;
                        fresh = true;
                    } else if (state == ST_RINGEND) {
                        ++kr;
                        q0 = 0;
                        state = ST_BATCH;
                        const float lb = (float)(2 * kr - 2) - D;
                        if (kr > kmax || (lb > 0.f && lb * lb > gb2)) state = ST_FINAL;
                    } else {
                        break;                           // ST_FINAL, drained: the group is done
                    }
                }
                if (fresh) {                             // inclusive scan of the list lengths over the lanes
                    uint32_t incl = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += up;
                    }
                    total = __shfl_sync(0xffffffffu, incl, 31);
                    excl = incl - cnt;
                    pos = 0;
                    const uint32_t owners = __ballot_sync(0xffffffffu, cnt != 0u);
                    single = (owners & (owners - 1u)) == 0u && owners ? __ffs(owners) - 1 : -1;   // one lane owns the whole list: no search
                }
            }
            if (pos < total) {
                // ---- the next 32 entries of the list: this lane's entry box against each voxel of the group with that voxel's bound
                const uint32_t q = pos + lane;
                pos += 32;
                // owner = the last lane whose exclusive offset is <= q (a lane with an empty list shares its successor's
                // offset and loses to it)
                int own = single;
                if (single < 0) {
                    own = 0;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        const int c2 = own + d;
                        const uint32_t ex = __shfl_sync(0xffffffffu, excl, c2 & 31);
                        if (ex <= q) own = c2;
                    }
                }
                const uint32_t ost = __shfl_sync(0xffffffffu, st, own), oex = __shfl_sync(0xffffffffu, excl, own);
                if (q < total) {
                    const uint4 en = __ldg(Em + ost + (q - oex));
                    tid = en.w;
                    ++n_tests;
                    const float2 a = unpack_h2(en.x), b = unpack_h2(en.y), c = unpack_h2(en.z);   // lo.x lo.y | lo.z hi.x | hi.y hi.z
                    // squared gap per axis for each node coordinate of the group on that axis
                    float ex[GX], ey[GY], ez[GZ];
#pragma unroll
                    for (int h = 0; h < GX; ++h) { const float gx = x0 + (float)h, t = fmaxf(0.f, fmaxf(a.x - gx, gx - b.y)); ex[h] = t * t; }
#pragma unroll
                    for (int h = 0; h < GY; ++h) { const float gy = y0 + (float)h, t = fmaxf(0.f, fmaxf(a.y - gy, gy - c.x)); ey[h] = t * t; }
#pragma unroll
                    for (int h = 0; h < GZ; ++h) { const float gz = z0 + (float)h, t = fmaxf(0.f, fmaxf(b.x - gz, gz - c.y)); ez[h] = t * t; }
                    // (x, y) partial sums once, the voxels' bounds four at a time
                    float exy[GX * GY];
#pragma unroll
                    for (int h = 0; h < GX * GY; ++h) exy[h] = ex[h % GX] + ey[h / GX];
#pragma unroll
                    for (int v4 = 0; v4 < NV; v4 += 4) {
                        float b4[4];
                        asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4[0]), "=f"(b4[1]), "=f"(b4[2]), "=f"(b4[3])
                                     : "r"((uint32_t)__cvta_generic_to_shared(&s_vb[warp][v4])) : "memory");
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (exy[(v4 + u) % (GX * GY)] + ez[(v4 + u) / (GX * GY)] <= b4[u]) vm |= 1u << (v4 + u);
                    }
                }
            }
            // ---- pairs of this batch to the queue, one pair per lane and round (a lane's lowest remaining voxel); whichever
            // stage has a full batch (or anything at all when draining, after the last round) runs, the later stage first so
            // that no queue overflows; each stage exists once
            bool um = __any_sync(0xffffffffu, vm != 0u);
            if (!drain && !um) continue;
            for (;;) {
                if (um) {
                    const bool mine = vm != 0u;
                    const uint32_t m = __ballot_sync(0xffffffffu, mine);
                    if (mine) {
                        const int vx = __ffs(vm) - 1;
                        vm &= vm - 1u;
                        pairs[pqn + __popc(m & lt)] = (tid << 5) | (uint32_t)vx;
                    }
                    pqn += __popc(m);
                    um = __any_sync(0xffffffffu, vm != 0u);
                }
                const bool last = !um && drain;
                if (pqn < 32 && !last) { if (!um) break; continue; }
                for (;;) {
                    const int act = qn >= 32 ? 2 : pqn >= 32 ? 1 : !last ? 0 : pqn > 0 ? 1 : (final && qn > 0) ? 2 : 0;
                    if (act == 0) break;
                    __syncwarp();
                    if (act == 1) {
                        // stage 1 on up to 32 (voxel, triangle) pairs, one pair per lane
                        const int c = min(pqn, 32);
                        pqn -= c;
                        bool cand = false, flagged = false;
                        uint32_t pr = 0;
                        float lo2 = 0.f;
                        int o = 0;
                        if (lane < c) {
                            pr = pairs[pqn + lane];
                            o = pr & 31u;
                            const V3 go = node_pos(kCell * cx + o % GX, kCell * cy + (o / GX) % GY, kCell * cz + o / (GX * GY), P);
                            float e2;
                            const float a2 = fast_dist2(go, Fm + 5 * (int64_t)(pr >> 5), flagged, e2);
                            ++n_pairs;
                            lo2 = a2 - e2;
                            if (!flagged && a2 + e2 < __uint_as_float(ubv[o])) atomicMin(&s_ub[warp][o], __float_as_uint(a2 + e2));   // non-negative floats order like uints
                        }
                        __syncwarp();                    // the batch's own upper bounds already count for its filter
                        if (lane < c) cand = flagged || !(lo2 > fmaf(__uint_as_float(ubv[o]), 1.002f, kappa));
                        const uint32_t cm = __ballot_sync(0xffffffffu, cand);
                        if (cand) qu[qn + __popc(cm & lt)] = pr;
                        qn += __popc(cm);
                        __syncwarp();
                        load_bounds();
                    } else {
                        // stage 2 on up to 32 queued pairs: the bit-faithful function decides
                        const int c = min(qn, 32);
                        qn -= c;
                        if (lane < c) {
                            const uint32_t e = qu[qn + lane];
                            const int o = e & 31u;
                            const V3 go = node_pos(kCell * cx + o % GX, kCell * cy + (o / GX) % GY, kCell * cz + o / (GX * GY), P);
                            const float d = point_triangle_distance(go, load_trirec(Tm + 4 * (int64_t)(e >> 5)));
                            ++n_faithful;
                            atomicMin(&s_best[warp][o], ((u64)__float_as_uint(d) << 32) | (e >> 5));
                        }
                    }
                }
                if (!um) break;
            }
        }
        __syncwarp();
        if (valid) {
            const u64 kb = s_best[warp][lane];
            key[(int64_t)mesh * nvox + (i + (int64_t)ni * (j + (int64_t)nj * k))] =
                kb == key0 ? (((u64)__float_as_uint(far) << 32) | 0xffffffffull) : kb;
        }
        __syncwarp();
    }
    if (counter) {
        if (n_faithful) atomicAdd(counter + 0, n_faithful);
        if (n_pairs) atomicAdd(counter + 2, n_pairs);
        if (n_tests) atomicAdd(counter + 3, n_tests);
    }
}

}  // namespace exact

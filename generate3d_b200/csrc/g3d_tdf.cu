// g3d_tdf.cu -- the distance-field path on sm_100a (both modes), behind g3d_tdf_batch.
//
// Faithful mode reproduces make_level_set3 (src/dfgen/makelevelset3.cpp:118-183) bit for bit -- phi AND
// closest_tri -- for a batch of meshes, with the reference's sequential loops re-expressed as
// order-independent parallel steps and with every evaluation skipped that provably cannot change the result:
//
//  tdf_prep      one thread / triangle: gathers the three vertices into a 64-byte record (vertices + the
//                per-triangle invariants of the distance function, computed with the reference's rounded
//                operation order), the double-precision grid coordinates, the band box
//                [int(min)-b, int(max)+b+1] (cpp:131-137) and the x-ray intersection counts (cpp:147-160).
//                Only the PARITY of the counts is ever used (cpp:178), so they are kept as XOR-ed bits.
//  tdf_init      the band initialisation. The reference's serial "if (d < phi) {phi = d; closest = t}" over
//                t = 0..T-1 ends at (min d, lowest t attaining it); atomicMin on the 64-bit key
//                (float_bits(d) << 32 | t) is exactly that, in any order. A node farther from the triangle's
//                bounding box than its current phi cannot win and is dropped before the evaluation (9 in 10
//                are); survivors are compacted and evaluated 32 at a time. Two kernels: tiled (one CTA per
//                16^3 block of a mesh, keys of the block in shared memory; batches of small grids) and scatter
//                (one warp per triangle, keys in global memory; large grids, small batches).
//  tdf_sweep     one CTA (or one thread-block cluster) / mesh, the 2 x 8 Gauss-Seidel sweeps (cpp:57-80,
//                163-172). Within one sweep node (a,b,c) (sweep-relative) reads only nodes of smaller a+b+c,
//                so the nodes of one anti-diagonal level are independent: 16 x (ni+nj+nk-5) synchronous
//                wavefront steps instead of 16 x (n-1)^3 serial ones. See the kernels for the exact shortcuts
//                (duplicate candidates, unchanged / met neighbours, band-box flags, untouched levels; on grids
//                of at most 65,536 nodes the second pass runs in tdf_sweep_bm_kernel: activity bits, speculation).
//  tdf_finalize  one warp / grid row: running parity along i (cpp:174-182), x out_scale (main.cpp:119), and
//                the fused epilogues the pipeline applies later on the CPU: |sdf| <= 1 occupancy
//                (vox2mesh/main.cpp:92), clamp at trunc (dataset_loader.py:137-139), sign*log(|x|+1)
//                (losses.py:7-11).
//
// Exact mode (g3d_tdf_exact.cuh) shares prep and finalize and replaces init + sweeps by a culled brute force.
//
// HBM layout (workspace): tri records [T_total] 64 B; band boxes [T_total] 16 B; keys u64 per mesh in 4x4x4 bricks
// (hi = phi bits, lo = sweep change code | band-box flags | closest_tri, see kTriBits; KeyMap); parity bits u32 [n_mesh, ceil(nvox/32)];
// sweep order table (shared by all meshes); small grids: 17 change bitmaps per mesh + one of the grid-face nodes.
#include <cuda_fp16.h>
#include <limits.h>
#include <stdlib.h>
#include <type_traits>
#include <algorithm>

#include "g3d_internal.h"
#include "g3d_ptd.cuh"

namespace g3d {

namespace {

typedef unsigned long long u64;

// Low word of a faithful-mode key: change code (4 bits: 0 = band init, s + 1 = sweep s, the last sweep shares 15 with the
// one before -- nothing reads codes after it) | band-box flags (6 bits, see tdf_band_flags_kernel) | closest_tri (22 bits,
// all ones = none). Exact-mode keys carry the plain triangle index.
constexpr int kTriBits = 22, kFlagShift = 22, kCodeShift = 28;
constexpr uint32_t kTriMask = (1u << kTriBits) - 1u;
constexpr uint32_t kTriMaskExact = 0x7ffffffu;
constexpr uint32_t kCodeMask = 0xfu << kCodeShift;
__host__ __device__ inline uint32_t sweep_code(int s) { return (uint32_t)(s + 1 < 15 ? s + 1 : 15); }

struct TdfWs {
    float4*   tri;          // 4 x float4 per triangle: TriRec (g3d_ptd.cuh)
    int4*     box;          // i0|i1<<16, j0|j1<<16, k0|k1<<16, mesh
    u64*      key;
    uint32_t* parity;
    uint32_t* order;        // packed a | b<<10 | c<<20, grouped by level
    int32_t*  level_start;  // [nlev + 1]
    uint32_t* chg;          // [n_mesh][17][act_words]: nodes changed under each code (sweep activity bitmaps), small grids only
    uint32_t* bnd;          // [act_words]: nodes on a grid face
    uint32_t* gact;         // cluster sweeps: [n_mesh][act_words] activity bits, [n_mesh][nlev + 4] level counters, [n_mesh] changes
    uint32_t* glvcnt;
    int32_t*  gnchg;
    int       act_words;    // 0: the grid is too large for the sweep's activity bitmap
    size_t    bytes;
};

// The second-pass sweep kernel keeps one bit per grid node: in shared memory when that is at most 8 KB (grids of up to
// 65,536 nodes, any batch), in global memory for the small batches of larger grids that sweep with one cluster per mesh.
constexpr int kClusterMeshes = 37;              // launch_sweep gives a mesh a cluster only when n_mesh * 4 <= SMs (148)
inline bool act_in_smem(const g3d_tdf_params& p) { return (int64_t)p.ni * p.nj * p.nk <= 65536; }
inline int act_words_of(const g3d_tdf_params& p, int32_t n_mesh)
{
    const int64_t nvox = (int64_t)p.ni * p.nj * p.nk;
    return (act_in_smem(p) || (n_mesh <= kClusterMeshes && nvox < ((int64_t)1 << 31))) ? (int)((nvox + 31) / 32) : 0;
}

__host__ __device__ inline int64_t nvox_of(const g3d_tdf_params& p) { return (int64_t)p.ni * p.nj * p.nk; }

// Faithful-mode keys are stored in 4x4x4 bricks (512 B), x fastest inside a brick, so that one 128-byte line is a
// 4x4 (i, j) patch of one k-slice. A sweep touches a node together with its 7 upwind neighbours (a 2x2x2 cube) and
// visits the grid by anti-diagonal levels: in the linear layout a line (16 nodes along i) stays live for 18 levels
// and the lines in use per mesh (~70 KB) times the meshes resident on an SM overflow the L1; a patch line is live
// for 9 levels, the lanes of a warp share lines, and the live set (~25 KB per mesh) fits.
// index = kx(i) + ky(j) + kz(k), separable so that the 8 indices of a cube cost six terms and eight sums.
struct KeyMap { int sbj, sbk; int64_t stride; };    // key steps of one brick in j and in k; keys per mesh
__host__ __device__ inline KeyMap keymap_of(const g3d_tdf_params& p)
{
    const int nbi = (p.ni + 3) >> 2, nbj = (p.nj + 3) >> 2, nbk = (p.nk + 3) >> 2;
    return { 64 * nbi, 64 * nbi * nbj, (int64_t)64 * nbi * nbj * nbk };
}
__device__ __forceinline__ int kx(int i) { return ((i >> 2) << 6) | (i & 3); }
__device__ __forceinline__ int ky(int j, int sbj) { return (j >> 2) * sbj + ((j & 3) << 2); }
__device__ __forceinline__ int kz(int k, int sbk) { return (k >> 2) * sbk + ((k & 3) << 4); }
inline int nlev_of(const g3d_tdf_params& p)
{
    int l = (p.ni - 2) + (p.nj - 2) + (p.nk - 2) + 1;
    return (p.ni < 2 || p.nj < 2 || p.nk < 2) ? 0 : l;
}

TdfWs carve(void* base, int32_t n_mesh, int64_t total_tris, const g3d_tdf_params& p)
{
    TdfWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (char*)base + o; };
    int64_t nvox = nvox_of(p);
    int64_t pw = (nvox + 31) / 32;
    int64_t nsweep = (int64_t)(p.ni - 1) * (p.nj - 1) * (p.nk - 1);
    if (nsweep < 0) nsweep = 0;
    w.tri = (float4*)take((size_t)total_tris * 64);
    w.box = (int4*)take((size_t)total_tris * 16);
    w.key = (u64*)take((size_t)n_mesh * keymap_of(p).stride * 8);       // bricked, padded to whole bricks
    w.parity = (uint32_t*)take((size_t)n_mesh * pw * 4);
    w.order = (uint32_t*)take((size_t)nsweep * 4);
    w.level_start = (int32_t*)take((size_t)(nlev_of(p) + 2) * 4);
    w.act_words = act_words_of(p, n_mesh);
    w.chg = (uint32_t*)take((size_t)n_mesh * 17 * w.act_words * 4);
    w.bnd = (uint32_t*)take((size_t)w.act_words * 4);
    const bool gl = w.act_words > 0 && n_mesh <= kClusterMeshes;     // (also for small grids: tests force clusters onto them)
    w.gact = (uint32_t*)take(gl ? (size_t)n_mesh * w.act_words * 4 : 0);
    w.glvcnt = (uint32_t*)take(gl ? (size_t)n_mesh * (nlev_of(p) + 4) * 4 : 0);
    w.gnchg = (int32_t*)take(gl ? (size_t)n_mesh * 4 : 0);
    w.bytes = off;
    return w;
}

// ------------------------------------------------------------------ fill ----
__global__ void tdf_fill_kernel(u64* key, int64_t n_key, u64 init, uint32_t* parity, int64_t n_par,
                                int32_t* status, int32_t n_mesh, uint32_t* occ_bits, int64_t n_occ)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = t0; i < n_key; i += stride) key[i] = init;
    for (int64_t i = t0; i < n_par; i += stride) parity[i] = 0u;
    if (occ_bits) for (int64_t i = t0; i < n_occ; i += stride) occ_bits[i] = 0u;
    if (status) for (int64_t i = t0; i < n_mesh; i += stride) status[i] = 0;
}

__device__ __forceinline__ V3 node_pos(int i, int j, int k, const g3d_tdf_params& P)
{
    // Vec3f gx(i*dx+origin[0], ...) (cpp:139): int -> float, one mul, one add
    return { add(mul((float)i, P.dx), P.origin[0]), add(mul((float)j, P.dx), P.origin[1]),
             add(mul((float)k, P.dx), P.origin[2]) };
}

#include "g3d_tdf_exact.cuh"

using exact::prefetch_l1;

// ------------------------------------------------------------------ prep ----
__device__ __forceinline__ double dmin(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ int clampi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }

// makelevelset3.cpp:84-94, without FMA contraction
__device__ __forceinline__ int orientation2d(double x1, double y1, double x2, double y2, double& area2)
{
    area2 = __dsub_rn(__dmul_rn(y1, x2), __dmul_rn(x1, y2));
    if (area2 > 0) return 1;
    if (area2 < 0) return -1;
    if (y2 > y1) return 1;
    if (y2 < y1) return -1;
    if (x1 > x2) return 1;
    if (x1 < x2) return -1;
    return 0;
}

// makelevelset3.cpp:98-116
__device__ __forceinline__ bool point_in_triangle_2d(double x0, double y0, double x1, double y1,
                                                     double x2, double y2, double x3, double y3,
                                                     double& a, double& b, double& c)
{
    x1 = __dsub_rn(x1, x0); x2 = __dsub_rn(x2, x0); x3 = __dsub_rn(x3, x0);
    y1 = __dsub_rn(y1, y0); y2 = __dsub_rn(y2, y0); y3 = __dsub_rn(y3, y0);
    int sa = orientation2d(x2, y2, x3, y3, a);
    if (sa == 0) return false;
    int sb = orientation2d(x3, y3, x1, y1, b);
    if (sb != sa) return false;
    int sc = orientation2d(x1, y1, x2, y2, c);
    if (sc != sa) return false;
    double sum = __dadd_rn(__dadd_rn(a, b), c);
    a = __ddiv_rn(a, sum);
    b = __ddiv_rn(b, sum);
    c = __ddiv_rn(c, sum);
    return true;
}

__device__ __forceinline__ int mesh_of(const int64_t* off, int n_mesh, int64_t t)
{
    int lo = 0, hi = n_mesh;                       // off[lo] <= t < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= t) lo = mid; else hi = mid;
    }
    return lo;
}

// IT = uint32_t, or uint16_t for the host path's compact index arrays (every mesh below 65,536 vertices)
template <typename IT, int MINB>
__global__ void __launch_bounds__(256, MINB)
tdf_prep_kernel(const float* __restrict__ verts, const int64_t* __restrict__ vert_off,
                const IT* __restrict__ tris, const int64_t* __restrict__ tri_off,
                int64_t t_first, int64_t total_tris, int n_mesh, g3d_tdf_params P, float4* __restrict__ tri_out,
                int4* __restrict__ box_out, uint32_t* __restrict__ parity, int32_t* status,
                float4* __restrict__ fast_out, uint4* __restrict__ tmp_out, uint32_t* __restrict__ cell_count,
                uint32_t* __restrict__ dil)
{
    int64_t t = t_first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_tris) return;
    int m = mesh_of(tri_off, n_mesh, t);
    int64_t v0 = vert_off[m], nv = vert_off[m + 1] - v0;
    uint32_t ip = tris[3 * t], iq = tris[3 * t + 1], ir = tris[3 * t + 2];
    // closest_tri is the mesh-local index in 22 bits (27 in exact mode): triangles of one mesh beyond that cannot be represented
    const bool too_many = t - tri_off[m] >= (int64_t)(P.mode == G3D_MODE_EXACT ? kTriMaskExact : kTriMask);
    if (too_many || ip >= nv || iq >= nv || ir >= nv) {
        // undefined behaviour in the reference (out-of-bounds read); flag the mesh
        // and neutralise the triangle: NaN vertices never win a `<` and an empty box
        if (status) atomicOr(status + m, too_many ? G3D_MESH_TOO_MANY_TRIS : G3D_MESH_BAD_INDEX);
        float qnan = __int_as_float(0x7fc00000);
        tri_out[4 * t] = tri_out[4 * t + 1] = tri_out[4 * t + 2] = tri_out[4 * t + 3] = make_float4(qnan, qnan, qnan, qnan);
        if (box_out) box_out[t] = make_int4(1, 1, 1, m);        // lo=1 > hi=0 on every axis
        if (tmp_out) tmp_out[t] = make_uint4(0u, 0u, 0u, exact::kNoCell);   // exact mode: not part of the cell index
        if (fast_out) for (int q = 0; q < 5; ++q) fast_out[5 * t + q] = make_float4(qnan, qnan, qnan, qnan);
        return;
    }
    const float* pp = verts + 3 * (v0 + ip);
    const float* pq = verts + 3 * (v0 + iq);
    const float* pr = verts + 3 * (v0 + ir);
    float xp[3] = { pp[0], pp[1], pp[2] }, xq[3] = { pq[0], pq[1], pq[2] }, xr[3] = { pr[0], pr[1], pr[2] };
    store_trirec(tri_out + 4 * t, make_trirec({ xp[0], xp[1], xp[2] }, { xq[0], xq[1], xq[2] }, { xr[0], xr[1], xr[2] }));
    if (fast_out) exact::store_fastrec(fast_out + 5 * t, { xp[0], xp[1], xp[2] }, { xq[0], xq[1], xq[2] }, { xr[0], xr[1], xr[2] });
    const int dims[3] = { P.ni, P.nj, P.nk };
    double fp[3], fq[3], fr[3];
    int lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {                  // cpp:131-137
        double o = (double)P.origin[a], h = (double)P.dx;
        fp[a] = __ddiv_rn(__dsub_rn((double)xp[a], o), h);
        fq[a] = __ddiv_rn(__dsub_rn((double)xq[a], o), h);
        fr[a] = __ddiv_rn(__dsub_rn((double)xr[a], o), h);
        double mn = dmin(fp[a], dmin(fq[a], fr[a])), mx = dmax(fp[a], dmax(fq[a], fr[a]));
        long long l = (long long)x86_d2i(mn) - P.exact_band;
        long long u = (long long)x86_d2i(mx) + P.exact_band + 1;
        lo[a] = (int)(l < 0 ? 0 : (l > dims[a] - 1 ? dims[a] - 1 : l));
        hi[a] = (int)(u < 0 ? 0 : (u > dims[a] - 1 ? dims[a] - 1 : u));
    }
    if (box_out) box_out[t] = make_int4(lo[0] | (hi[0] << 16), lo[1] | (hi[1] << 16), lo[2] | (hi[2] << 16), m);
    if (tmp_out) {
        // exact mode: the triangle joins the mesh's cell index. A non-finite vertex gives NaN distances, which never
        // win in the reference: such a triangle is left out altogether.
        const float chk = xp[0] + xp[1] + xp[2] + xq[0] + xq[1] + xq[2] + xr[0] + xr[1] + xr[2];
        if (fabsf(chk) < 3.0e38f) exact::index_triangle(fp, fq, fr, m, t, P, tmp_out, cell_count, dil);
        else tmp_out[t] = make_uint4(0u, 0u, 0u, exact::kNoCell);
    }

    // cpp:147-160: x-ray intersection parity
    int j0 = clampi(x86_d2i(ceil(dmin(fp[1], dmin(fq[1], fr[1])))), 0, P.nj - 1);
    int j1 = clampi(x86_d2i(floor(dmax(fp[1], dmax(fq[1], fr[1])))), 0, P.nj - 1);
    int k0 = clampi(x86_d2i(ceil(dmin(fp[2], dmin(fq[2], fr[2])))), 0, P.nk - 1);
    int k1 = clampi(x86_d2i(floor(dmax(fp[2], dmax(fq[2], fr[2])))), 0, P.nk - 1);
    int64_t nvox = nvox_of(P);
    uint32_t* par = parity + (int64_t)m * ((nvox + 31) / 32);
    for (int k = k0; k <= k1; ++k)
        for (int j = j0; j <= j1; ++j) {
            double a, b, c;
            if (point_in_triangle_2d((double)j, (double)k, fp[1], fp[2], fq[1], fq[2], fr[1], fr[2], a, b, c)) {
                double fi = __dadd_rn(__dadd_rn(__dmul_rn(a, fp[0]), __dmul_rn(b, fq[0])), __dmul_rn(c, fr[0]));
                int iv = x86_d2i(ceil(fi));
                int bin = iv < 0 ? 0 : iv;
                if (bin < P.ni) {
                    int64_t idx = bin + (int64_t)P.ni * (j + (int64_t)P.nj * k);
                    atomicXor(par + (idx >> 5), 1u << (idx & 31));
                }
            }
        }
}

// ------------------------------------------------------------------ init ----

// One warp per triangle walks the triangle's band box 32 nodes at a time. A node farther from the
// triangle's bounding box than its current phi cannot win (nor tie) the `<`, so it is dropped
// before the expensive evaluation (about 9 in 10 are); the survivors of successive triangles are
// collected in a warp-private queue and evaluated 32 at a time, one (node, triangle) pair per lane.
// The margin (0.1 % on the squared distance) dwarfs the rounding of the float distance function;
// reading a stale, larger phi only makes the test less aggressive.
__device__ __forceinline__ void init_eval_item(uint32_t t, uint32_t pos, const float4* __restrict__ tri,
                                               const int4* __restrict__ box, const int64_t* __restrict__ tri_off,
                                               const g3d_tdf_params& P, u64* key, int64_t nvox, float far)
{
    const int i = pos & 1023u, j = (pos >> 10) & 1023u, k = pos >> 20;
    const int mesh = __ldg(box + t).w;
    const float d = point_triangle_distance(node_pos(i, j, k, P), load_trirec(tri + 4 * (int64_t)t));
    const KeyMap km = keymap_of(P);
    u64* kp = key + (int64_t)mesh * km.stride + (kx(i) + ky(j, km.sbj) + kz(k, km.sbk));
    const float cur = __uint_as_float((uint32_t)(__ldca(kp) >> 32));
    // d < far mirrors the very first `d < phi`; d <= cur lets an equal distance with a LOWER
    // triangle index through, as the serial loop would
    if (d < far && d <= cur) atomicMin(kp, ((u64)__float_as_uint(d) << 32) | (uint32_t)(t - tri_off[mesh]));
}

__global__ void __launch_bounds__(256)
tdf_init_kernel(const float4* __restrict__ tri, const int4* __restrict__ box,
                const int64_t* __restrict__ tri_off, int64_t t_first, int64_t total_tris, g3d_tdf_params P,
                u64* __restrict__ key, float far, u64* counter)
{
    __shared__ uint32_t s_qt[8][64], s_qp[8][64];
    u64 n_eval = 0;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* qt = s_qt[wib];
    uint32_t* qp = s_qp[wib];
    int qn = 0;                                     // queue fill, warp-uniform
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t nvox = nvox_of(P);
    const KeyMap km = keymap_of(P);
    for (int64_t t = t_first + warp; t < total_tris; t += nwarp) {
        if (lane < 2 && t + nwarp < total_tris) {    // the next triangle's box and record are two dependent loads away
            if (lane == 0) prefetch_l1(box + t + nwarp);
            else prefetch_l1(tri + 4 * (t + nwarp));
        }
        int4 b = __ldg(box + t);
        int i0 = b.x & 0xffff, i1 = b.x >> 16, j0 = b.y & 0xffff, j1 = b.y >> 16, k0 = b.z & 0xffff, k1 = b.z >> 16;
        int bi = i1 - i0 + 1, bj = j1 - j0 + 1, bk = k1 - k0 + 1;
        if (bi <= 0 || bj <= 0 || bk <= 0) continue;
        const float4 v1 = __ldg(tri + 4 * t), v2 = __ldg(tri + 4 * t + 1), v3 = __ldg(tri + 4 * t + 2);
        const float lox = fminf(v1.x, fminf(v2.x, v3.x)), hix = fmaxf(v1.x, fmaxf(v2.x, v3.x));
        const float loy = fminf(v1.y, fminf(v2.y, v3.y)), hiy = fmaxf(v1.y, fmaxf(v2.y, v3.y));
        const float loz = fminf(v1.z, fminf(v2.z, v3.z)), hiz = fmaxf(v1.z, fmaxf(v2.z, v3.z));
        // a non-finite vertex gives NaN distances, which never win: nothing to do for this triangle
        if (!(fabsf(v1.x + v1.y + v1.z + v2.x + v2.y + v2.z + v3.x + v3.y + v3.z) < 3.0e38f)) continue;
        const u64* K = key + (int64_t)b.w * km.stride;
        const int n = bi * bj * bk;
        const float rbi = 1.0f / (float)bi, rbj = 1.0f / (float)bj;
        for (int q0 = 0; q0 < n; q0 += 32) {
            const int q = q0 + lane;
            bool pass = false;
            uint32_t pos = 0;
            if (q < n) {
                // q = i' + bi*(j' + bj*k'). For q < 2^20 the float product (q + 0.5) * fl(1/bi) is within
                // q * 2^-22 < 0.5/bi of the true quotient, whose fraction is at least 0.5/bi away from an
                // integer, so truncation gives the exact quotient without an integer division.
                int r, k;
                if (n <= (1 << 20)) {
                    r = (int)(((float)q + 0.5f) * rbi);
                    k = (int)(((float)r + 0.5f) * rbj);
                } else {
                    r = q / bi;
                    k = r / bj;
                }
                const int i = i0 + (q - r * bi);
                const int j = j0 + (r - k * bj);
                k += k0;
                const V3 gx = node_pos(i, j, k, P);
                const float cur = __uint_as_float((uint32_t)(__ldca(K + (kx(i) + ky(j, km.sbj) + kz(k, km.sbk))) >> 32));
                const float ex = fmaxf(0.f, fmaxf(lox - gx.x, gx.x - hix));
                const float ey = fmaxf(0.f, fmaxf(loy - gx.y, gx.y - hiy));
                const float ez = fmaxf(0.f, fmaxf(loz - gx.z, gx.z - hiz));
                pass = !(fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > fmaf(cur * cur, 1.001f, 1e-12f));
                pos = (uint32_t)i | ((uint32_t)j << 10) | ((uint32_t)k << 20);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                if (pass) {
                    const int w = qn + __popc(m & ((1u << lane) - 1u));
                    qt[w] = (uint32_t)t;
                    qp[w] = pos;
                }
                qn += __popc(m);
                if (qn >= 32) {
                    __syncwarp();
                    qn -= 32;
                    init_eval_item(qt[qn + lane], qp[qn + lane], tri, box, tri_off, P, key, nvox, far);
                    ++n_eval;
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();
    if (lane < qn) {
        init_eval_item(qt[lane], qp[lane], tri, box, tri_off, P, key, nvox, far);
        ++n_eval;
    }
    if (counter && n_eval) atomicAdd(counter, n_eval);
}

// Tiled variant for batches of small grids: one CTA owns a kTile^3 block of one mesh's grid and keeps its keys in
// shared memory. Its warps scan the mesh's band boxes 32 at a time and collect the boxes that reach into the block;
// the cull reads and the atomicMin go to shared memory: no L2 round trip in front of every test, a fresher phi (so
// about half as many pairs survive), and the block is written out once with plain stores. A triangle whose band
// spans several blocks is visited by each of them, for its own nodes only.
//   Typical band boxes hold 27-64 nodes, so the per-triangle overhead of a warp-wide walk would dominate. Small
// boxes (after clipping) are therefore walked one per LANE: 32 pending triangles at a time, each lane stepping
// through the rows of its own box, one cull test per lane and step. Large boxes keep the
// warp-wide walk. Survivors of both go to the same warp-private queue and are evaluated 32 at a time.
// Band-box flags of the node (i, j, k) for a triangle with (clamped) band box b: does the box also hold the node's
// neighbour at i+1, i-1, j+1, j-1, k+1, k-1 (bits 0..5). The band init evaluated the triangle at every node of its box, so
// a sweep that finds the triangle at a neighbour of such a node need not evaluate it again there: the flags of the axes
// the two nodes differ in say whether that node is inside the box (on the other axes it shares the flagged node's
// coordinate, which is inside).
__device__ __forceinline__ uint32_t band_flags(int4 b, int i, int j, int k)
{
    const int i0 = b.x & 0xffff, i1 = b.x >> 16, j0 = b.y & 0xffff, j1 = b.y >> 16, k0 = b.z & 0xffff, k1 = b.z >> 16;
    return (i + 1 <= i1 ? 1u : 0u) | (i - 1 >= i0 ? 2u : 0u) | (j + 1 <= j1 ? 4u : 0u) | (j - 1 >= j0 ? 8u : 0u) |
           (k + 1 <= k1 ? 16u : 0u) | (k - 1 >= k0 ? 32u : 0u);
}

constexpr int kTile = 16;
constexpr int kLaneI = 8, kLaneRows = 64;            // lane-walked boxes: at most 8 wide in i, 64 (j, k) rows

// One (node, triangle) pair per lane against the block's keys in shared memory. Kept out of line: the walk loops
// below reach it from many unrolled sites and the evaluation is several hundred instructions long.
__device__ __noinline__ void tile_eval_item(const float4* __restrict__ TR, u64* s_key, uint32_t t, uint32_t pos, int I0,
                                            int J0, int K0, float dx, float ox, float oy, float oz, float far)
{
    const int i = pos & 1023u, j = (pos >> 10) & 1023u, k = pos >> 20;
    const V3 gx = { add(mul((float)i, dx), ox), add(mul((float)j, dx), oy), add(mul((float)k, dx), oz) };   // node_pos
    const float d = point_triangle_distance(gx, load_trirec(TR + 4 * (int64_t)t));
    const int e = (i - I0) + kTile * ((j - J0) + kTile * (k - K0));
    const float cur = __uint_as_float(((const volatile uint32_t*)s_key)[2 * e + 1]);
    // d < far mirrors the very first `d < phi`; d <= cur lets an equal distance with a LOWER index through
    if (d < far && d <= cur) atomicMin(s_key + e, ((u64)__float_as_uint(d) << 32) | t);
}

__global__ void __launch_bounds__(256, 4)
tdf_init_tiled_kernel(const float4* __restrict__ tri, const int4* __restrict__ box, const int64_t* __restrict__ tri_off,
                      int mesh0, int nbx, int nby, int nbz, g3d_tdf_params P, u64* __restrict__ key, float far,
                      u64* counter)
{
    __shared__ u64 s_key[kTile * kTile * kTile];
    __shared__ uint32_t s_qt[8][64], s_qp[8][64], s_pend[8][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int per = nbx * nby * nbz;
    const int mesh = mesh0 + (int)(blockIdx.x / (unsigned)per);
    const int rb = (int)(blockIdx.x % (unsigned)per);
    const int I0 = (rb % nbx) * kTile, J0 = ((rb / nbx) % nby) * kTile, K0 = (rb / (nbx * nby)) * kTile;
    const int I1 = min(I0 + kTile, P.ni) - 1, J1 = min(J0 + kTile, P.nj) - 1, K1 = min(K0 + kTile, P.nk) - 1;
    const u64 init = ((u64)__float_as_uint(far) << 32) | 0xffffffffull;
    for (int e = threadIdx.x; e < kTile * kTile * kTile; e += 256) s_key[e] = init;
    __syncthreads();
    const volatile uint32_t* s_w = (const volatile uint32_t*)s_key;
    uint32_t* qt = s_qt[wib];
    uint32_t* qp = s_qp[wib];
    uint32_t* pend = s_pend[wib];
    int qn = 0, pn = 0;                             // queue / pending fill, warp-uniform
    u64 n_eval = 0;
    const int64_t tb = tri_off[mesh];
    const int T = (int)(tri_off[mesh + 1] - tb);
    const float4* TR = tri + 4 * tb;
    const int4* BX = box + tb;

    auto eval_item = [&](uint32_t t, uint32_t pos) {
        tile_eval_item(TR, s_key, t, pos, I0, J0, K0, P.dx, P.origin[0], P.origin[1], P.origin[2], far);
    };
    // survivors of one step (warp-wide): compact into the queue, evaluate a full warp's worth when there is one
    auto push = [&](bool pass, uint32_t t, uint32_t pos) {
        const uint32_t m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            if (pass) {
                const int w = qn + __popc(m & lt);
                qt[w] = t;
                qp[w] = pos;
            }
            qn += __popc(m);
            if (qn >= 32) {
                __syncwarp();
                qn -= 32;
                eval_item(qt[qn + lane], qp[qn + lane]);
                ++n_eval;
                __syncwarp();
            }
        }
    };
    // squared distance from a node to the triangle's bounding box against the node's current phi
    auto cull_pass = [&](int i, int j, int k, float lox, float hix, float loy, float hiy, float loz, float hiz) {
        const V3 gx = node_pos(i, j, k, P);
        const float cur = __uint_as_float(s_w[2 * ((i - I0) + kTile * ((j - J0) + kTile * (k - K0))) + 1]);
        const float ex = fmaxf(0.f, fmaxf(lox - gx.x, gx.x - hix));
        const float ey = fmaxf(0.f, fmaxf(loy - gx.y, gx.y - hiy));
        const float ez = fmaxf(0.f, fmaxf(loz - gx.z, gx.z - hiz));
        return !(fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > fmaf(cur * cur, 1.001f, 1e-12f));
    };
    // one triangle per lane (t < 0: idle lane); boxes at most kLaneI nodes wide in i after clipping. The lane
    // steps through the (j, k) rows of its box; along a row only the i term of the box distance changes, and
    // those (at most kLaneI) squares are computed once per triangle.
    auto walk_lanes = [&](int t) {
        int i0 = I0, bi = 0, j0 = J0, j1 = J0 - 1, k0 = K0, nrow = 0;
        float ex2[kLaneI];
        float loy = 0.f, hiy = 0.f, loz = 0.f, hiz = 0.f;
#pragma unroll
        for (int a = 0; a < kLaneI; ++a) ex2[a] = 0.f;
        if (t >= 0) {
            const int4 b = __ldg(BX + t);
            i0 = max(b.x & 0xffff, I0);
            bi = min(b.x >> 16, I1) - i0 + 1;
            j0 = max(b.y & 0xffff, J0); j1 = min(b.y >> 16, J1);
            k0 = max(b.z & 0xffff, K0);
            const int k1 = min(b.z >> 16, K1);
            const float4 v1 = __ldg(TR + 4 * (int64_t)t), v2 = __ldg(TR + 4 * (int64_t)t + 1), v3 = __ldg(TR + 4 * (int64_t)t + 2);
            const float lox = fminf(v1.x, fminf(v2.x, v3.x)), hix = fmaxf(v1.x, fmaxf(v2.x, v3.x));
            loy = fminf(v1.y, fminf(v2.y, v3.y)); hiy = fmaxf(v1.y, fmaxf(v2.y, v3.y));
            loz = fminf(v1.z, fminf(v2.z, v3.z)); hiz = fmaxf(v1.z, fmaxf(v2.z, v3.z));
            nrow = (j1 - j0 + 1) * (k1 - k0 + 1);
            // a non-finite vertex gives NaN distances, which never win: nothing to do for this triangle
            if (!(fabsf(v1.x + v1.y + v1.z + v2.x + v2.y + v2.z + v3.x + v3.y + v3.z) < 3.0e38f)) nrow = 0;
#pragma unroll
            for (int a = 0; a < kLaneI; ++a) {
                const float gx = add(mul((float)(i0 + a), P.dx), P.origin[0]);
                const float ex = fmaxf(0.f, fmaxf(lox - gx, gx - hix));
                ex2[a] = ex * ex;
            }
        }
        const int rmax = __reduce_max_sync(0xffffffffu, nrow), bimax = __reduce_max_sync(0xffffffffu, bi);
        int j = j0, k = k0;
        for (int r = 0; r < rmax; ++r) {
            const bool rv = r < nrow;
            const float gy = add(mul((float)j, P.dx), P.origin[1]), gz = add(mul((float)k, P.dx), P.origin[2]);
            const float ey = fmaxf(0.f, fmaxf(loy - gy, gy - hiy)), ez = fmaxf(0.f, fmaxf(loz - gz, gz - hiz));
            const float eyz2 = fmaf(ey, ey, ez * ez);
            const int e = (i0 - I0) + kTile * ((j - J0) + kTile * (k - K0));
            const uint32_t pos = (uint32_t)i0 | ((uint32_t)j << 10) | ((uint32_t)k << 20);
            // the row's nodes that survive the cull as a bit mask per lane; only the positions at which some lane has a
            // survivor go through the (warp-wide) compaction
            uint32_t pm = 0;
#pragma unroll
            for (int a = 0; a < kLaneI; ++a) {
                if (rv && a < bi) {
                    const float cur = __uint_as_float(s_w[2 * (e + a) + 1]);
                    if (!(ex2[a] + eyz2 > fmaf(cur * cur, 1.001f, 1e-12f))) pm |= 1u << a;
                }
            }
            for (uint32_t um = __reduce_or_sync(0xffffffffu, pm); um; um &= um - 1u) {
                const int a = __ffs(um) - 1;
                push((pm >> a) & 1u, (uint32_t)t, pos + (uint32_t)a);
            }
            if (++j > j1) { j = j0; ++k; }
        }
    };
    // the whole warp on one (large) box
    auto walk_warp = [&](int t, int bxw, int byw, int bzw) {
        const int i0 = max(bxw & 0xffff, I0), i1 = min(bxw >> 16, I1), j0 = max(byw & 0xffff, J0),
                  j1 = min(byw >> 16, J1), k0 = max(bzw & 0xffff, K0), k1 = min(bzw >> 16, K1);
        const int bi = i1 - i0 + 1, bj = j1 - j0 + 1, bk = k1 - k0 + 1;
        const float4 v1 = __ldg(TR + 4 * (int64_t)t), v2 = __ldg(TR + 4 * (int64_t)t + 1), v3 = __ldg(TR + 4 * (int64_t)t + 2);
        const float lox = fminf(v1.x, fminf(v2.x, v3.x)), hix = fmaxf(v1.x, fmaxf(v2.x, v3.x));
        const float loy = fminf(v1.y, fminf(v2.y, v3.y)), hiy = fmaxf(v1.y, fmaxf(v2.y, v3.y));
        const float loz = fminf(v1.z, fminf(v2.z, v3.z)), hiz = fmaxf(v1.z, fmaxf(v2.z, v3.z));
        if (!(fabsf(v1.x + v1.y + v1.z + v2.x + v2.y + v2.z + v3.x + v3.y + v3.z) < 3.0e38f)) return;
        const int n = bi * bj * bk;                      // <= kTile^3
        const float rbi = 1.0f / (float)bi, rbj = 1.0f / (float)bj;
        for (int q0 = 0; q0 < n; q0 += 32) {
            const int q = q0 + lane;
            bool pass = false;
            uint32_t pos = 0;
            if (q < n) {
                const int r = (int)(((float)q + 0.5f) * rbi);     // exact for q < 2^20, see above
                int k = (int)(((float)r + 0.5f) * rbj);
                const int i = i0 + (q - r * bi);
                const int j = j0 + (r - k * bj);
                k += k0;
                pass = cull_pass(i, j, k, lox, hix, loy, hiy, loz, hiz);
                pos = (uint32_t)i | ((uint32_t)j << 10) | ((uint32_t)k << 20);
            }
            push(pass, (uint32_t)t, pos);
        }
    };

    for (int base = wib * 32; base < T; base += 8 * 32) {
        const int tl = base + lane;
        int4 b = make_int4(1, 1, 1, 0);
        bool small = false, large = false;
        if (tl < T) {
            b = __ldg(BX + tl);
            const int ci = min(b.x >> 16, I1) - max(b.x & 0xffff, I0) + 1, cj = min(b.y >> 16, J1) - max(b.y & 0xffff, J0) + 1,
                      ck = min(b.z >> 16, K1) - max(b.z & 0xffff, K0) + 1;
            if (ci > 0 && cj > 0 && ck > 0) {
                prefetch_l1(TR + 4 * (int64_t)tl);
                large = ci > kLaneI || cj * ck > kLaneRows;
                small = !large;
            }
        }
        uint32_t hm = __ballot_sync(0xffffffffu, large);
        while (hm) {
            const int src = __ffs(hm) - 1;
            hm &= hm - 1;
            walk_warp(base + src, __shfl_sync(0xffffffffu, b.x, src), __shfl_sync(0xffffffffu, b.y, src),
                      __shfl_sync(0xffffffffu, b.z, src));
        }
        const uint32_t sm = __ballot_sync(0xffffffffu, small);
        if (sm) {
            if (small) pend[pn + __popc(sm & lt)] = (uint32_t)tl;
            pn += __popc(sm);
            if (pn >= 32) {
                __syncwarp();
                pn -= 32;
                walk_lanes((int)pend[pn + lane]);
                __syncwarp();
            }
        }
    }
    __syncwarp();
    if (pn > 0) walk_lanes(lane < pn ? (int)pend[lane] : -1);
    __syncwarp();
    if (lane < qn) {
        eval_item(qt[lane], qp[lane]);
        ++n_eval;
    }
    __syncthreads();
    const KeyMap km = keymap_of(P);
    u64* K = key + (int64_t)mesh * km.stride;
    for (int e = threadIdx.x; e < kTile * kTile * kTile; e += 256) {
        const int i = I0 + (e % kTile), j = J0 + (e / kTile) % kTile, k = K0 + e / (kTile * kTile);
        if (i <= I1 && j <= J1 && k <= K1) {
            u64 kv = s_key[e];
            const uint32_t t = (uint32_t)kv;
            if (t != 0xffffffffu) kv |= (u64)band_flags(__ldg(BX + t), i, j, k) << kFlagShift;
            K[kx(i) + ky(j, km.sbj) + kz(k, km.sbk)] = kv;
        }
    }
    if (counter && n_eval) atomicAdd(counter, n_eval);
}

// ------------------------------------------------------- band-box flags ----
// The flags (see band_flags) for meshes [m0, m0 + n_mesh) after the scatter band init; the tiled one sets them as it
// writes its block out.
__global__ void __launch_bounds__(256)
tdf_band_flags_kernel(u64* __restrict__ key, const int4* __restrict__ box, const int64_t* __restrict__ tri_off, int m0,
                      int n_mesh, g3d_tdf_params P)
{
    const int64_t nvox = nvox_of(P);
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nvox * n_mesh) return;
    const int mesh = m0 + (int)(q / nvox);
    const int64_t v = q - (int64_t)(mesh - m0) * nvox;
    const int i = (int)(v % P.ni), j = (int)((v / P.ni) % P.nj), k = (int)(v / ((int64_t)P.ni * P.nj));
    const KeyMap km = keymap_of(P);
    uint32_t* lo = reinterpret_cast<uint32_t*>(key + (int64_t)mesh * km.stride + (kx(i) + ky(j, km.sbj) + kz(k, km.sbk)));
    const uint32_t w = *lo;
    if (w == 0xffffffffu) return;                       // no triangle
    *lo = w | (band_flags(__ldg(box + tri_off[mesh] + w), i, j, k) << kFlagShift);
}

// ----------------------------------------------------------- sweep order ----
// Nodes of the sweep lattice (ni-1)(nj-1)(nk-1) grouped by level a+b+c.
__global__ void __launch_bounds__(1024)
tdf_order_kernel(int na, int nb, int nc, int nlev, uint32_t* __restrict__ order, int32_t* __restrict__ level_start,
                 uint32_t* __restrict__ bnd, int act_words)
{
    // (the grid is (na+1) x (nb+1) x (nc+1) nodes) bit i + ni (j + nj k) of bnd: the node lies on a grid face
    // (every CTA builds the small tables for itself and takes a share of the two big loops)
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < act_words; w += gridDim.x * blockDim.x) {
        uint32_t m = 0u;
        const int p0 = 32 * w;
        int i = p0 % (na + 1), r = p0 / (na + 1), j = r % (nb + 1), k = r / (nb + 1);
        for (int b = 0; b < 32; ++b) {
            if (k <= nc && (i == 0 || i == na || j == 0 || j == nb || k == 0 || k == nc)) m |= 1u << b;
            if (++i > na) { i = 0; if (++j > nb) { j = 0; ++k; } }
        }
        bnd[w] = m;
    }
    // Closed form, no atomics: within a level nodes are sorted by (c, b). Neighbouring lanes of a sweep warp then
    // hold neighbouring grid rows, whose upwind rows coincide, so a warp's loads fall into few cache lines -- and
    // the order is the same on every run.
    //   cnt2(m) = #{(a, b): a + b = m};  F(x) = sum of cnt2 up to x;  nodes of level l in slabs c' < c: F(l) - F(l - c).
    extern __shared__ int32_t s_tab[];             // F[na + nb - 1], then level starts [nlev + 1]
    const int nm = na + nb - 1;
    int32_t* F = s_tab;
    int32_t* ls = s_tab + nm;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int m = 0; m < nm; ++m) {
            acc += min(min(m, na + nb - 2 - m), min(na, nb) - 1) + 1;
            F[m] = acc;
        }
        auto Fx = [&](int x) { return x < 0 ? 0 : F[min(x, nm - 1)]; };
        acc = 0;
        for (int l = 0; l < nlev; ++l) {
            ls[l] = acc;
            if (blockIdx.x == 0) level_start[l] = acc;
            acc += Fx(l) - Fx(l - nc);
        }
        ls[nlev] = acc;
        if (blockIdx.x == 0) level_start[nlev] = acc;
    }
    __syncthreads();
    auto Fx = [&](int x) { return x < 0 ? 0 : F[min(x, nm - 1)]; };
    const int64_t n = (int64_t)na * nb * nc;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int a = (int)(q % na), r = (int)(q / na);
        const int b = r % nb, c = r / nb;
        const int m = a + b, l = m + c;
        const int pos = ls[l] + Fx(l) - Fx(l - c) + (b - max(0, m - (na - 1)));
        order[pos] = (uint32_t)a | ((uint32_t)b << 10) | ((uint32_t)c << 20);
    }
}

// ----------------------------------------------------------------- sweep ----
// One CTA per mesh. Per level each thread owns one node and must replay, in the reference's
// order, "if (d(node, closest_tri(nb)) < phi) take it" for its 7 upwind neighbours. A few
// observations remove most of that work without changing a single bit of the result
// (d is a pure function of (node, triangle) and phi only ever decreases, so a comparison
// that failed once fails for ever, and a pair that was evaluated -- or skipped because it was
// known to fail -- leaves phi(node) <= d(node, triangle) behind for good):
//   1. a candidate equal to the node's own triangle, or to an earlier candidate of the same
//      visit, cannot pass the strict `<`;
//   2. a neighbour whose closest_tri has not changed since this node last examined it (in an
//      earlier sweep that had the same neighbour upwind) cannot pass either. The top 4 bits of a
//      key's low word record the sweep in which each node last changed; s_thr[s][f][r] is the last
//      earlier sweep in which slot r of sweep s pointed at the same absolute neighbour, for a node
//      lying on the grid faces f (such a node is not part of every sweep's lattice);
//   3. nor can a candidate equal to the triangle of ANY neighbour that falls under 2: the node met
//      that triangle when it examined that neighbour (sweep_need_mask);
//   4. a neighbour that still holds its band-init triangle, with this node inside that triangle's
//      band box, counts as met too: the band init evaluated the pair (band_flags, sweep_candidates);
//   5. the surviving (node, triangle) pairs are unevenly spread over the lanes, so each warp
//      pushes them into a private shared-memory queue, evaluates the queue densely (one pair
//      per lane), and the owners then replay the results in neighbour order.
// tests/sweep_rules_model.c states 1-4 (and the speculation of tdf_sweep_bm_kernel) on the CPU and is checked
// against the plain oracle without a GPU.
constexpr int kSweepQueue = 32 * 7;                 // worst case per warp and round
// (three rounds need a 196 KB shared-memory carve-out at 7 CTAs per SM and lose more L1 than they gain: 13.8 ms
// against 12.5 with two and 13.0 with one)
__host__ __device__ constexpr int sweep_rounds(int nt) { return nt <= 256 ? 2 : 1; }   // rounds of a level sharing one queue

__device__ __forceinline__ void sweep_dir(int s, int& di, int& dj, int& dk)
{
    const int q8 = s & 7;                           // cpp:163-172, pass 0 then pass 1
    di = (q8 & 1) ? -1 : 1;
    dj = (q8 == 0 || q8 == 2 || q8 == 5 || q8 == 7) ? 1 : -1;
    dk = (q8 == 0 || q8 == 3 || q8 == 4 || q8 == 7) ? 1 : -1;
}

// CL = true: one mesh per thread-block CLUSTER (large grids, small batches): the CTAs of a cluster split every
// level between them and meet at a hardware cluster barrier, whose release/acquire also orders the global-memory
// keys between the SMs involved.
__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }

// ---- candidate selection of a sweep visit (shared by both sweep kernels) ----
// Thresholds of the current sweep as words: thrw[f][r] = (threshold code + 1) << kCodeShift, so that "the neighbour has not
// changed since this node last examined it" is the unsigned comparison w < thrw (the change code sits in the top bits of a
// key's low word). ib[r]: the band-box flags a neighbour at slot r must carry for this node to lie inside the band box of
// the neighbour's triangle (the node is the neighbour's neighbour in the sweep direction on the axes of r + 1). Rebuilt by
// the first threads of the CTA at the start of every sweep (callers put a barrier behind it).
__device__ __forceinline__ void sweep_threshold_words(int s, const int8_t (*thr_s)[8], uint32_t (*thrw)[8], uint32_t* ib)
{
    if (threadIdx.x < 64) thrw[threadIdx.x >> 3][threadIdx.x & 7] = (uint32_t)((int)thr_s[threadIdx.x >> 3][threadIdx.x & 7] + 1) << kCodeShift;
    if (threadIdx.x >= 64 && threadIdx.x < 72) {
        int di, dj, dk;
        sweep_dir(s, di, dj, dk);
        const int a = (int)threadIdx.x - 63;            // 1..8 (8: padding)
        const uint32_t f = ((a & 1) ? (di > 0 ? 1u : 2u) : 0u) | ((a & 2) ? (dj > 0 ? 4u : 8u) : 0u) | ((a & 4) ? (dk > 0 ? 16u : 32u) : 0u);
        ib[a - 1] = a < 8 ? f << kFlagShift : 0xffffffffu;
    }
}
// Which of the 7 upwind neighbours (low words w[r] of their keys, thresholds tw[r]) need an evaluation; t[r] receives the
// triangle fields. cm: the candidates -- neighbours that hold a triangle the node has not met through them. A neighbour is
// "met" (xm) when it has not changed since the node last examined it, or when it still holds its band-init triangle and
// its flags say that this node lies inside that triangle's band box: the band init evaluated the pair.
// A candidate cannot pass the strict `<` when it is the node's own triangle (own: the triangle field of its key), an
// earlier candidate of the same visit, or the triangle of ANY met neighbour: the node met that triangle then (evaluated
// it, or skipped it for one of these very reasons), and phi has only decreased since. The last case costs nothing to
// detect -- the seven words are loaded anyway -- and removes four evaluations in ten (the same triangle reaching a node
// through a second neighbour).
__device__ __forceinline__ uint32_t sweep_candidates(const uint32_t (&w)[7], const uint32_t (&tw)[7], const uint32_t (&ib)[7],
                                                     uint32_t (&t)[7], uint32_t& xm)
{
    uint32_t cm = 0;
    xm = 0;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        t[r] = w[r] & kTriMask;
        const bool ex = w[r] < tw[r] || (w[r] & (ib[r] | kCodeMask)) == ib[r];
        xm |= ex ? 1u << r : 0u;
        cm |= (!ex && t[r] != kTriMask) ? 1u << r : 0u;
    }
    return cm;
}
__device__ __forceinline__ uint32_t sweep_need_mask(const uint32_t (&t)[7], uint32_t xm, uint32_t cm, uint32_t own)
{
    // one comparison per pair of slots: eq bit r = slots q < r hold the same triangle. Then slot r is dropped (an earlier
    // slot holds its triangle: a candidate evaluated first, or a neighbour examined before), and slot q is dropped when one
    // of its later twins is an examined neighbour. (Two empty slots compare equal too; they are no candidates anyway.)
    uint32_t drop = 0;
#pragma unroll
    for (int r = 0; r < 7; ++r) drop |= t[r] == own ? 1u << r : 0u;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        uint32_t eq = 0;
#pragma unroll
        for (int r = q + 1; r < 7; ++r) eq |= t[q] == t[r] ? 1u << r : 0u;
        drop |= eq;
        drop |= (eq & xm) ? 1u << q : 0u;
    }
    return cm & ~drop;
}

template <int NT, bool CL>
__global__ void __launch_bounds__(NT)
tdf_sweep_kernel(u64* key, const float4* __restrict__ tri, const int64_t* __restrict__ tri_off,
                 const uint32_t* __restrict__ order, const int32_t* __restrict__ level_start, int nlev,
                 int lev_words, g3d_tdf_params P, u64* counter, int s_begin, int s_end, uint32_t* chg, int act_words)
{
    u64 n_eval = 0;
    extern __shared__ uint32_t s_mem[];
    // s_thr[s][f][r]: code of the last sweep before s in which a node examined the neighbour that slot r of sweep s points
    // at (-1: never). f = the axes on which the node lies on a grid face (bit 0: x ...): such a node is only part of the
    // sweeps that run towards its face, so for it an earlier sweep counts only if it ran the same way on those axes too.
    // In all: the last sweep that agrees with s on the axes of (r + 1) | f.
    __shared__ __align__(8) int8_t s_thr[16][8][8];    // (r = 7: padding)
    __shared__ __align__(16) uint32_t s_thrw[8][8];    // the current sweep's thresholds as words, see sweep_threshold_words
    __shared__ __align__(16) uint32_t s_ib[8];
    int32_t* s_lev = reinterpret_cast<int32_t*>(s_mem);                         // [nlev + 1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int R = sweep_rounds(NT);
    uint32_t* q_tri = s_mem + lev_words + warp * (2 * R * kSweepQueue);
    uint32_t* q_pos = q_tri + R * kSweepQueue;
    // q_pos[e] is replaced by the pair's result (by the lane that read it). Shared memory is worth saving here:
    // what the CTAs of an SM do not take stays L1 (measured, 1,024 meshes: flat up to a 164 KB carve-out with the
    // bricked keys, +9 % at 196 KB, +100 % with the L1 squeezed to its minimum).
    // s_dirty[d][l + 3]: newest change code among the nodes whose own level in direction d is l (-3 <= l < nlev; nodes on
    // the grid faces lie outside some sweep lattices). A node of level l is upwind of nodes of levels l+1 .. l+3, so level
    // L has nothing new upwind when the three entries below it are old: one store per direction and change, three loads
    // per level.
    const int dstride = nlev + 3;
    uint8_t* s_dirty = reinterpret_cast<uint8_t*>(s_mem + lev_words + (NT / 32) * (2 * R * kSweepQueue));   // [8][nlev + 3]
    uint32_t* s_lut = s_mem + lev_words + (NT / 32) * (2 * R * kSweepQueue) + (8 * dstride + 3) / 4;       // [na + nb + nc]
    for (int l = threadIdx.x; l <= nlev; l += NT) s_lev[l] = level_start[l];
    for (int l = threadIdx.x; l < 8 * dstride; l += NT) s_dirty[l] = 0;
    for (int e = threadIdx.x; e < 16 * 8 * 8; e += NT) {
        const int r = e & 7, f = (e >> 3) & 7, s = e >> 6, mask = (r + 1) | f;
        int d[3], thr = -1;
        sweep_dir(s, d[0], d[1], d[2]);
        for (int sp = 0; sp < s; ++sp) {
            int e3[3];
            sweep_dir(sp, e3[0], e3[1], e3[2]);
            bool ok = true;
            for (int ax = 0; ax < 3; ++ax)
                if (((mask >> ax) & 1) && e3[ax] != d[ax]) ok = false;
            if (ok) thr = sp + 1;
        }
        s_thr[s][f][r] = (int8_t)thr;
    }
    __syncthreads();
    const int csize = CL ? (int)cluster_nctarank() : 1, crank = CL ? (int)cluster_ctarank() : 0;
    const int mesh = blockIdx.x / csize;
    const KeyMap km = keymap_of(P);
    u64* K = key + (int64_t)mesh * km.stride;
    // low word of each key = change code << 27 | closest_tri (kTriNone when the node has no triangle yet): one
    // 4-byte load tells both whether a neighbour needs to be examined and which triangle it holds
    const uint32_t* Klo = (const uint32_t*)K;
    const float4* T = tri + 4 * tri_off[mesh];
    const int ni = P.ni, nj = P.nj, nk = P.nk, na = ni - 1, nb = nj - 1, nc = nk - 1;

    // chg (optional): per-code change bitmaps of the mesh for tdf_sweep_bm_kernel, which takes over the later sweeps
    uint32_t* const Cm = chg ? chg + (int64_t)mesh * 17 * act_words : nullptr;
    for (int s = s_begin; s < s_end; ++s) {
        int di, dj, dk;
        sweep_dir(s, di, dj, dk);
        const int tmin = s_thr[s][7][6];            // the oldest threshold of the sweep: the last sweep in the same direction
        const volatile uint8_t* dirty = s_dirty + (s & 7) * dstride;
        sweep_threshold_words(s, s_thr[s], s_thrw, s_ib);    // (the last level of the previous sweep ended with a barrier)
        // per sweep-relative coordinate: key offset of the node | upwind neighbour in another brick << 30 | grid face << 31
        for (int e = threadIdx.x; e < na + nb + nc; e += NT) {
            const int ax = e < na ? 0 : (e < na + nb ? 1 : 2), rel = e - (ax == 0 ? 0 : (ax == 1 ? na : na + nb));
            const int d = ax == 0 ? di : (ax == 1 ? dj : dk), nn = ax == 0 ? ni : (ax == 1 ? nj : nk);
            const int v = d > 0 ? 1 + rel : nn - 2 - rel;
            const int off = ax == 0 ? kx(v) : (ax == 1 ? ky(v, km.sbj) : kz(v, km.sbk));
            const bool cross = d > 0 ? (v & 3) == 0 : (v & 3) == 3, face = v == 0 || v == nn - 1;
            s_lut[e] = (uint32_t)off | (cross ? 0x40000000u : 0u) | (face ? 0x80000000u : 0u);
        }
        // key steps to the upwind neighbour inside a brick / across bricks, and the node of sweep-relative coordinate 0
        const int3 step = make_int3(di, 4 * dj, 16 * dk), stepc = make_int3(61 * di, (km.sbj - 12) * dj, (km.sbk - 48) * dk);
        const int3 base = make_int3(di > 0 ? 1 : ni - 2, dj > 0 ? 1 : nj - 2, dk > 0 ? 1 : nk - 2);
        if (CL) cluster_barrier(); else __syncthreads();
        for (int L = 0; L < nlev; ++L) {
            // nothing upwind of this whole level has changed since it was last examined: every comparison the
            // reference would make here is known to fail. (Cluster variant: the table is per CTA, no skipping.)
            if (!CL && max(max((int)dirty[L], (int)dirty[L + 1]), (int)dirty[L + 2]) <= tmin) continue;
            const int beg = s_lev[L], end = s_lev[L + 1];
            // node phase of one round: candidates of this lane's node go to the queue from slot `qbase` on; returns
            // the number of pairs the warp queued (warp-uniform)
            auto node_phase = [&](int q, int qbase, uint32_t& need, uint32_t& pos, float& phi, int& ct, int& n, int& off) -> int {
                int c0 = -1, c1 = -1, c2 = -1, c3 = -1, c4 = -1, c5 = -1, c6 = -1;
                need = 0; pos = 0; phi = 0.f; ct = -1; n = 0; off = 0;
                if (q < end) {
                    // the node is known by its sweep-relative coordinates (pos = a | b << 10 | c << 20); key offsets, brick
                    // crossings and grid faces per coordinate come from the sweep's tables
                    pos = __ldg(order + q);
                    const uint32_t lx = s_lut[pos & 1023u], ly = s_lut[na + ((pos >> 10) & 1023u)], lz = s_lut[na + nb + (pos >> 20)];
                    // the node and its 7 upwind neighbours (in the order of cpp:72-78) in the bricked key layout
                    const int x0 = (int)(lx & 0x3fffffffu), y0 = (int)(ly & 0x3fffffffu), z0 = (int)(lz & 0x3fffffffu);
                    const int x1 = x0 - ((lx & 0x40000000u) ? stepc.x : step.x), y1 = y0 - ((ly & 0x40000000u) ? stepc.y : step.y),
                              z1 = z0 - ((lz & 0x40000000u) ? stepc.z : step.z);
                    n = x0 + y0 + z0;
                    const int n0 = x1 + y0 + z0, n1 = x0 + y1 + z0, n2 = x1 + y1 + z0, n3 = x0 + y0 + z1,
                              n4 = x1 + y0 + z1, n5 = x0 + y1 + z1, n6 = x1 + y1 + z1;
                    const int face = (int)((lx >> 31) | ((ly >> 31) << 1) | ((lz >> 31) << 2));
                    const uint4 tA = *reinterpret_cast<const uint4*>(&s_thrw[face][0]), tB = *reinterpret_cast<const uint4*>(&s_thrw[face][4]);
                    const uint32_t tw[7] = { tA.x, tA.y, tA.z, tA.w, tB.x, tB.y, tB.z };
                    const uint32_t wd[7] = { Klo[2 * n0], Klo[2 * n1], Klo[2 * n2], Klo[2 * n3], Klo[2 * n4], Klo[2 * n5], Klo[2 * n6] };
                    const uint4 iA = *reinterpret_cast<const uint4*>(&s_ib[0]), iB = *reinterpret_cast<const uint4*>(&s_ib[4]);
                    const uint32_t ib[7] = { iA.x, iA.y, iA.z, iA.w, iB.x, iB.y, iB.z };
                    uint32_t tt[7], xm;
                    const uint32_t cm = sweep_candidates(wd, tw, ib, tt, xm);
                    c0 = (int)tt[0]; c1 = (int)tt[1]; c2 = (int)tt[2]; c3 = (int)tt[3]; c4 = (int)tt[4]; c5 = (int)tt[5]; c6 = (int)tt[6];
                    if (cm) {                                            // some neighbour holds a triangle worth trying
                        u64 kv = K[n];
                        phi = __uint_as_float((uint32_t)(kv >> 32));
                        const uint32_t own = (uint32_t)kv & kTriMask;
                        ct = own == kTriMask ? -1 : (int)own;
                        need = sweep_need_mask(tt, xm, cm, own);
                    }
                }
                if (!__any_sync(0xffffffffu, need != 0)) return 0;
                // warp-exclusive prefix of the per-lane work counts
                const int cnt = __popc(need);
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                off = qbase + incl - cnt;
                int w = off;
                if (need & 1u)  { q_tri[w] = c0; q_pos[w] = pos; ++w; }
                if (need & 2u)  { q_tri[w] = c1; q_pos[w] = pos; ++w; }
                if (need & 4u)  { q_tri[w] = c2; q_pos[w] = pos; ++w; }
                if (need & 8u)  { q_tri[w] = c3; q_pos[w] = pos; ++w; }
                if (need & 16u) { q_tri[w] = c4; q_pos[w] = pos; ++w; }
                if (need & 32u) { q_tri[w] = c5; q_pos[w] = pos; ++w; }
                if (need & 64u) { q_tri[w] = c6; q_pos[w] = pos; ++w; }
                return __shfl_sync(0xffffffffu, incl, 31);
            };
            // the owner replays the results of its pairs in neighbour order (they lie in that order from `off` on)
            auto apply = [&](uint32_t need, uint32_t pos, float phi, int ct, int n, int off) {
                if (!need) return;
                bool changed = false;
                for (int w = off; need; need &= need - 1, ++w) {
                    const float d = __uint_as_float(q_pos[w]);
                    if (d < phi) { phi = d; ct = (int)q_tri[w]; changed = true; }
                }
                if (changed) {
                    const uint8_t code = (uint8_t)(s + 1);
                    K[n] = ((u64)__float_as_uint(phi) << 32) | (sweep_code(s) << kCodeShift) | (uint32_t)ct;
                    if (!CL) {
                        // stamp the node's own level in every direction (codes only grow and all writers of a sweep
                        // write the same value). No first-pass sweep can skip a level (each has neighbours nobody has
                        // looked at yet), so the stamps are only kept when this kernel runs second-pass sweeps too.
                        const int pi = base.x + di * (int)(pos & 1023u), pj = base.y + dj * (int)((pos >> 10) & 1023u), pk = base.z + dk * (int)(pos >> 20);
                        if (s_end > 8) {
#pragma unroll
                            for (int dq = 0; dq < 8; ++dq) {
                                int ei, ej, ek;
                                sweep_dir(dq, ei, ej, ek);
                                const int lv = (ei > 0 ? pi - 1 : ni - 2 - pi) + (ej > 0 ? pj - 1 : nj - 2 - pj) +
                                               (ek > 0 ? pk - 1 : nk - 2 - pk);
                                s_dirty[dq * dstride + lv + 3] = code;
                            }
                        }
                    }
                    if (Cm) {
                        const int pi = base.x + di * (int)(pos & 1023u), pj = base.y + dj * (int)((pos >> 10) & 1023u), pk = base.z + dk * (int)(pos >> 20);
                        const int p = pi + ni * (pj + nj * pk);
                        atomicOr(Cm + (s + 1) * act_words + (p >> 5), 1u << (p & 31));
                    }
                }
            };
            // The nodes of a level are independent: a warp runs the node phase of two of its rounds back to back
            // (their loads overlap, their pairs share one queue), evaluates the queue densely, then applies.
            for (int qb = beg + crank * NT + warp * 32; qb < end; qb += R * NT * csize) {
                uint32_t needA, posA, needB = 0, posB = 0;
                float phiA, phiB = 0.f;
                int ctA, nA, offA, ctB = -1, nB = 0, offB = 0;
                int W = node_phase(qb + lane, 0, needA, posA, phiA, ctA, nA, offA);
                if (R == 2 && qb + NT * csize < end) W += node_phase(qb + NT * csize + lane, W, needB, posB, phiB, ctB, nB, offB);
                if (W == 0) continue;
                __syncwarp();
                for (int e = lane; e < W; e += 32) {
                    uint32_t pp = q_pos[e];
                    V3 gx = node_pos(base.x + di * (int)(pp & 1023u), base.y + dj * (int)((pp >> 10) & 1023u), base.z + dk * (int)(pp >> 20), P);
                    q_pos[e] = __float_as_uint(point_triangle_distance(gx, load_trirec(T + 4 * (int64_t)q_tri[e])));
                    ++n_eval;
                }
                __syncwarp();
                apply(needA, posA, phiA, ctA, nA, offA);
                if (R == 2) apply(needB, posB, phiB, ctB, nB, offB);
                __syncwarp();                       // the queue is free again
            }
            if (CL) cluster_barrier(); else __syncthreads();
        }
    }
    if (counter && n_eval) atomicAdd(counter + 1, n_eval);
}

// ------------------------------------------------- sweep, small grids ----
// The sweep kernel for grids of at most 65,536 nodes (one CTA per mesh): the same three shortcuts as tdf_sweep_kernel
// below, plus two that need one bit per grid node in shared memory.
//   4. Activity bits. s_act (linear index i + ni (j + nj k)) marks the nodes that may have a candidate in this sweep: a
//      SUPERSET of the nodes whose candidate test can pass (the exact test still runs for every node visited), built at
//      the start of a sweep from per-code change bitmaps and extended by every change of the sweep itself. The nodes of
//      a level are scanned against it and only the marked ones, compacted, run the node phase.
//   5. Speculation. Once sweeps change few nodes, a sweep first visits ALL its marked nodes at once, level order ignored,
//      against the state the sweep started from, and only notes which of them would change. A node whose upwind
//      neighbours do not change during the sweep meets exactly that state when the reference visits it, so "would not
//      change" is final for it. What is left -- the would-change nodes and, as changes happen, the nodes downwind of
//      them -- is then swept level by level as usual; that is a handful of nodes, and levels without any are skipped.
//      The second pass of a converged field costs one parallel visit per sweep instead of ~90 dependent level steps.

// A node changed in sweep s: stamp its level in every direction, put the change on record under this sweep's code and
// mark the 7 nodes downwind of it, which have a new candidate in this very sweep (they lie on later levels; lvcnt counts
// the marked nodes per level).
template <bool SP>
__device__ __forceinline__ void sweep_record_change(uint32_t pos, int s, int di, int dj, int dk, int ni, int nj, int nk,
                                                    int dstride, uint8_t* s_dirty, uint32_t* chg_code, uint32_t* s_act,
                                                    uint32_t* s_lvcnt)
{
    const uint8_t code = (uint8_t)(s + 1);
    const int pi = (int)(pos & 1023u), pj = (int)((pos >> 10) & 1023u), pk = (int)(pos >> 20);
    int lv_s = 0;
#pragma unroll
    for (int dq = 0; dq < 8; ++dq) {            // (codes only grow and all writers of a sweep write the same value)
        int ei, ej, ek;
        sweep_dir(dq, ei, ej, ek);
        const int lv = (ei > 0 ? pi - 1 : ni - 2 - pi) + (ej > 0 ? pj - 1 : nj - 2 - pj) + (ek > 0 ? pk - 1 : nk - 2 - pk);
        s_dirty[dq * dstride + lv + 3] = code;
        if (dq == (s & 7)) lv_s = lv;
    }
    const int p = pi + ni * (pj + nj * pk);
    atomicOr(chg_code + (p >> 5), 1u << (p & 31));
    if (!SP) return;                            // (no activity bits in a plain sweep)
    const bool ox = (unsigned)(pi + di) < (unsigned)ni, oy = (unsigned)(pj + dj) < (unsigned)nj,
               oz = (unsigned)(pk + dk) < (unsigned)nk;
    const int sx = di, sy = dj * ni, sz = dk * ni * nj;
#pragma unroll
    for (int a = 1; a < 8; ++a) {
        const bool in = (!(a & 1) || ox) && (!(a & 2) || oy) && (!(a & 4) || oz);
        if (in) {
            const int pp = p + ((a & 1) ? sx : 0) + ((a & 2) ? sy : 0) + ((a & 4) ? sz : 0);
            const uint32_t bit = 1u << (pp & 31);
            const uint32_t old = atomicOr(&s_act[pp >> 5], bit);
            if (!(old & bit)) atomicAdd(&s_lvcnt[lv_s + ((a & 1) + ((a >> 1) & 1) + (a >> 2))], 1u);
        }
    }
}

// code-0 change bitmap of a mesh: the nodes that hold a triangle after the band init (all threads of the CTA)
__device__ __noinline__ void sweep_init_bitmap(const uint32_t* Klo, KeyMap km, int ni, int nj, int nk, uint32_t* Cm, int act_words)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nvox = ni * nj * nk;
    for (int w = warp; w < act_words; w += (int)(blockDim.x >> 5)) {
        const int p = 32 * w + lane;
        bool has = false;
        if (p < nvox) {
            const int i = p % ni, r = p / ni, j = r % nj, k = r / nj;
            has = (Klo[2 * (kx(i) + ky(j, km.sbj) + kz(k, km.sbk))] & kTriMask) != kTriMask;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, has);
        if (lane == 0) Cm[w] = m;
    }
    __syncthreads();
}

// s_act for sweep s (all threads of the CTA). Node n has a candidate at slot r iff its neighbour m = n - d*r holds a
// triangle and changed in a sweep after the one in which n last examined it, i.e. iff m is in the union of the change
// bitmaps of the codes above the slot's threshold; that union, shifted by the slot's offset in the linear index, marks the
// nodes n. Interior nodes use their exact thresholds (thw: the seven bytes of s_thr[s][0]); nodes on a grid face, whose
// thresholds are older, use the oldest one of the sweep (tmin). Shifts that run across a row end only set extra bits.
template <bool CL>
__device__ __noinline__ void sweep_build_activity(int s, int tmin, uint2 thw, const uint32_t* Cm, int act_words,
                                                  const uint32_t* __restrict__ bnd, uint32_t* s_act, int ni, int nj)
{
    int di, dj, dk;
    sweep_dir(s, di, dj, dk);
    // (CL: the bitmap lives in global memory and the cluster's CTAs share the words)
    const int t0 = (CL ? (int)cluster_ctarank() * (int)blockDim.x : 0) + (int)threadIdx.x;
    const int tn = (CL ? (int)cluster_nctarank() : 1) * (int)blockDim.x;
    for (int w = t0; w < act_words; w += tn) s_act[w] = 0u;
    if (CL) cluster_barrier(); else __syncthreads();
    int th[7];
#pragma unroll
    for (int r = 0; r < 7; ++r) th[r] = (int)(int8_t)((r < 4 ? thw.x : thw.y) >> (8 * (r & 3)));
    const int nc = s - tmin;                           // codes tmin + 1 .. s: at most 8 (first pass: s + 1; second: 7)
    for (int w = t0; w < act_words; w += tn) {
        uint32_t cw[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) cw[q] = q < nc ? __ldcg(Cm + (s - q) * act_words + w) : 0u;
        uint32_t acc = 0u, v[7] = { 0u, 0u, 0u, 0u, 0u, 0u, 0u };
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            acc |= cw[q];
#pragma unroll
            for (int r = 0; r < 7; ++r) if (th[r] + 1 == s - q) v[r] = acc;
        }
        if (acc == 0u) continue;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const int a = r + 1;
            const int delta = ((a & 1) ? di : 0) + ((a & 2) ? dj * ni : 0) + ((a & 4) ? dk * ni * nj : 0);
            const int bit = 32 * w + delta;
            const int wd = bit >> 5, sh = bit & 31;             // arithmetic shift: floor for negative bit positions
            const uint32_t lo_i = v[r] << sh, lo_b = acc << sh;
            const uint32_t hi_i = sh ? v[r] >> (32 - sh) : 0u, hi_b = sh ? acc >> (32 - sh) : 0u;
            if (wd >= 0 && wd < act_words && lo_b) {
                const uint32_t val = lo_i | (lo_b & __ldg(bnd + wd));
                if (val) atomicOr(&s_act[wd], val);
            }
            if (wd + 1 >= 0 && wd + 1 < act_words && hi_b) {
                const uint32_t val = hi_i | (hi_b & __ldg(bnd + wd + 1));
                if (val) atomicOr(&s_act[wd + 1], val);
            }
        }
    }
    if (CL) cluster_barrier(); else __syncthreads();
}

// CL: one mesh per thread-block cluster (large grids / small batches, second-pass sweeps only): the activity bitmap, the
// level counters and the change count live in global memory, the CTAs share every loop and meet at cluster barriers.
template <int NT, bool CL>
__global__ void __launch_bounds__(NT, NT == 128 ? 7 : (NT == 256 ? 3 : 1))
tdf_sweep_bm_kernel(u64* key, const float4* __restrict__ tri, const int64_t* __restrict__ tri_off,
                    const uint32_t* __restrict__ order, const int32_t* __restrict__ level_start, int nlev,
                    int lev_words, g3d_tdf_params P, u64* counter, uint32_t* chg, const uint32_t* __restrict__ bnd,
                    int act_words, int s_begin, int s_end, int spec_div, uint32_t* gact, uint32_t* glvcnt, int* gnchg)
{
    auto csync = [] { if (CL) cluster_barrier(); else __syncthreads(); };
    const int csize = CL ? (int)cluster_nctarank() : 1, crank = CL ? (int)cluster_ctarank() : 0;
    u64 n_eval = 0;
    extern __shared__ uint32_t s_mem[];
    __shared__ __align__(8) int8_t s_thr[16][8][8];    // as in tdf_sweep_kernel
    __shared__ __align__(16) uint32_t s_thrw[8][8];
    __shared__ __align__(16) uint32_t s_ib[8];
    __shared__ int s_nchg_cta;                         // nodes changed in the sweep so far (one CTA per mesh)
    int32_t* s_lev = reinterpret_cast<int32_t*>(s_mem);                         // [nlev + 1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    constexpr int R = sweep_rounds(NT);
    uint32_t* q_tri = s_mem + lev_words + warp * (2 * R * kSweepQueue);
    uint32_t* q_pos = q_tri + R * kSweepQueue;
    const int dstride = nlev + 3;
    uint8_t* s_dirty = reinterpret_cast<uint8_t*>(s_mem + lev_words + (NT / 32) * (2 * R * kSweepQueue));   // [8][nlev + 3]
    const int mesh = blockIdx.x / csize;
    uint32_t* s_act = CL ? gact + (size_t)mesh * act_words
                         : s_mem + lev_words + (NT / 32) * (2 * R * kSweepQueue) + (8 * dstride + 3) / 4;               // [act_words]
    uint32_t* s_lvcnt = CL ? glvcnt + (size_t)mesh * (nlev + 4) : s_act + act_words;   // [nlev + 4]: marked nodes per level (speculative sweeps)
    int* const s_nchg = CL ? gnchg + mesh : &s_nchg_cta;
    uint32_t* s_list = q_tri;                          // the nodes waiting for a batch; the queue is idle while they collect
    for (int l = threadIdx.x; l <= nlev; l += NT) s_lev[l] = level_start[l];
    // (taking over from tdf_sweep_kernel after sweep s_begin - 1: its level stamps are gone, so a plain sweep here skips no level)
    for (int l = threadIdx.x; l < 8 * dstride; l += NT) s_dirty[l] = s_begin ? 255 : 0;
    for (int e = threadIdx.x; e < 16 * 8 * 8; e += NT) {
        const int r = e & 7, f = (e >> 3) & 7, s = e >> 6, mask = (r + 1) | f;
        int d[3], thr = -1;
        sweep_dir(s, d[0], d[1], d[2]);
        for (int sp = 0; sp < s; ++sp) {
            int e3[3];
            sweep_dir(sp, e3[0], e3[1], e3[2]);
            bool ok = true;
            for (int ax = 0; ax < 3; ++ax)
                if (((mask >> ax) & 1) && e3[ax] != d[ax]) ok = false;
            if (ok) thr = sp + 1;
        }
        s_thr[s][f][r] = (int8_t)thr;
    }
    if (threadIdx.x == 0 && crank == 0) *s_nchg = s_begin ? 0 : 0x7fffffff;
    const KeyMap km = keymap_of(P);
    u64* K = key + (int64_t)mesh * km.stride;
    const uint32_t* Klo = (const uint32_t*)K;
    const float4* T = tri + 4 * tri_off[mesh];
    const int ni = P.ni, nj = P.nj, nk = P.nk;
    const int n_lattice = (ni - 1) * (nj - 1) * (nk - 1);
    const float rcp_ni = 1.0f / (float)ni, rcp_nj = 1.0f / (float)nj;
    // change bitmaps of this mesh, one per code (0 = the band init, s + 1 = sweep s), zero at launch
    uint32_t* Cm = chg + (int64_t)mesh * 17 * act_words;
    csync();
    if (s_begin == 0) {                             // (never with a cluster: those take over after the first pass)
        sweep_init_bitmap(Klo, km, ni, nj, nk, Cm, act_words);
    } else {                                        // the nodes sweep s_begin - 1 changed
        int c = 0;
        for (int w = crank * NT + threadIdx.x; w < act_words; w += NT * csize) c += __popc(__ldcg(Cm + s_begin * act_words + w));
        if (c) atomicAdd(s_nchg, c);
    }

    // One sweep. SP = false: the plain level-synchronous sweep of tdf_sweep_kernel (every node of every level that is not
    // skipped whole), for sweeps that still change many nodes. SP = true: activity bits + speculation.
    auto sweep = [&](auto sp_tag, const int s) {
        constexpr bool SP = decltype(sp_tag)::value;
        int di, dj, dk;
        sweep_dir(s, di, dj, dk);
        const int tmin = s_thr[s][7][6];            // the oldest threshold of the sweep: the last sweep in the same direction
        const volatile uint8_t* dirty = s_dirty + (s & 7) * dstride;
        if (SP) {
            for (int l = crank * NT + threadIdx.x; l < nlev + 4; l += NT * csize) s_lvcnt[l] = 0u;
            if (s <= 7) {
                // first pass: each sweep brings neighbours no sweep has looked at before, so about every node that has a
                // neighbour with a triangle is active: all bits on, no bitmap arithmetic
                for (int w = crank * NT + threadIdx.x; w < act_words; w += NT * csize) s_act[w] = 0xffffffffu;
                csync();
            } else {
                sweep_build_activity<CL>(s, tmin, *reinterpret_cast<const uint2*>(&s_thr[s][0][0]), Cm, act_words, bnd, s_act, ni, nj);
            }
        }
        uint32_t* const chg_code = Cm + (s + 1) * act_words;
        int n_changed = 0;                          // this thread's nodes changed in the sweep
        for (int pass = SP ? 0 : 1; pass < 2; ++pass) {
        const bool flat = SP && pass == 0;          // the speculative visit of every marked node, levels ignored
        for (int L = 0; L < (flat ? 1 : nlev); ++L) {
            int beg, end;
            if (flat) {
                beg = 0; end = s_lev[nlev];
            } else {
                if (SP) {
                    // on to the next level that holds a marked node, 32 counters per look (a level-by-level look costs a
                    // cluster a global-memory round trip per level)
                    bool found = false;
                    while (L < nlev) {
                        const uint32_t c = L + lane < nlev ? ((const volatile uint32_t*)s_lvcnt)[L + lane] : 0u;
                        const uint32_t m = __ballot_sync(0xffffffffu, c != 0u);
                        if (m) { L += __ffs(m) - 1; found = true; break; }
                        L += 32;
                    }
                    if (!found) break;
                }
                // nothing upwind of this whole level has changed since it was last examined
                else if (!CL && max(max((int)dirty[L], (int)dirty[L + 1]), (int)dirty[L + 2]) <= tmin) continue;
                beg = s_lev[L]; end = s_lev[L + 1];
            }
            // node phase of one round: candidates of this lane's node go to the queue from slot `qbase` on; returns
            // the number of pairs the warp queued (warp-uniform)
            auto node_phase = [&](bool valid, uint32_t pos, int qbase, uint32_t& need, float& phi, int& ct, int& n, int& off) -> int {
                int c0 = -1, c1 = -1, c2 = -1, c3 = -1, c4 = -1, c5 = -1, c6 = -1;
                need = 0; phi = 0.f; ct = -1; n = 0; off = 0;
                if (valid) {
                    const int i = (int)(pos & 1023u), j = (int)((pos >> 10) & 1023u), k = (int)(pos >> 20);
                    const int x0 = kx(i), x1 = kx(i - di), y0 = ky(j, km.sbj), y1 = ky(j - dj, km.sbj),
                              z0 = kz(k, km.sbk), z1 = kz(k - dk, km.sbk);
                    n = x0 + y0 + z0;
                    const int n0 = x1 + y0 + z0, n1 = x0 + y1 + z0, n2 = x1 + y1 + z0, n3 = x0 + y0 + z1,
                              n4 = x1 + y0 + z1, n5 = x0 + y1 + z1, n6 = x1 + y1 + z1;
                    const int face = ((i == 0 || i == ni - 1) ? 1 : 0) | ((j == 0 || j == nj - 1) ? 2 : 0) | ((k == 0 || k == nk - 1) ? 4 : 0);
                    const uint4 tA = *reinterpret_cast<const uint4*>(&s_thrw[face][0]), tB = *reinterpret_cast<const uint4*>(&s_thrw[face][4]);
                    const uint32_t tw[7] = { tA.x, tA.y, tA.z, tA.w, tB.x, tB.y, tB.z };
                    const uint32_t wd[7] = { Klo[2 * n0], Klo[2 * n1], Klo[2 * n2], Klo[2 * n3], Klo[2 * n4], Klo[2 * n5], Klo[2 * n6] };
                    const uint4 iA = *reinterpret_cast<const uint4*>(&s_ib[0]), iB = *reinterpret_cast<const uint4*>(&s_ib[4]);
                    const uint32_t ib[7] = { iA.x, iA.y, iA.z, iA.w, iB.x, iB.y, iB.z };
                    uint32_t tt[7], xm;
                    const uint32_t cm = sweep_candidates(wd, tw, ib, tt, xm);
                    c0 = (int)tt[0]; c1 = (int)tt[1]; c2 = (int)tt[2]; c3 = (int)tt[3]; c4 = (int)tt[4]; c5 = (int)tt[5]; c6 = (int)tt[6];
                    if (cm) {                                            // some neighbour holds a triangle worth trying
                        u64 kv = K[n];
                        phi = __uint_as_float((uint32_t)(kv >> 32));
                        const uint32_t own = (uint32_t)kv & kTriMask;
                        ct = own == kTriMask ? -1 : (int)own;
                        need = sweep_need_mask(tt, xm, cm, own);
                    }
                }
                if (!__any_sync(0xffffffffu, need != 0)) return 0;
                const int cnt = __popc(need);
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                off = qbase + incl - cnt;
                int w = off;
                if (need & 1u)  { q_tri[w] = c0; q_pos[w] = pos; ++w; }
                if (need & 2u)  { q_tri[w] = c1; q_pos[w] = pos; ++w; }
                if (need & 4u)  { q_tri[w] = c2; q_pos[w] = pos; ++w; }
                if (need & 8u)  { q_tri[w] = c3; q_pos[w] = pos; ++w; }
                if (need & 16u) { q_tri[w] = c4; q_pos[w] = pos; ++w; }
                if (need & 32u) { q_tri[w] = c5; q_pos[w] = pos; ++w; }
                if (need & 64u) { q_tri[w] = c6; q_pos[w] = pos; ++w; }
                return __shfl_sync(0xffffffffu, incl, 31);
            };
            // the owner replays the results of its pairs in neighbour order (they lie in that order from `off` on).
            // Speculative visit: nothing is stored; a node that would not change loses its mark, one that would keeps it
            // and is counted on its level.
            auto apply = [&](bool valid, uint32_t need, uint32_t pos, float phi, int ct, int n, int off) {
                bool changed = false;
                for (int w = off; need; need &= need - 1, ++w) {
                    const float d = __uint_as_float(q_pos[w]);
                    if (d < phi) { phi = d; ct = (int)q_tri[w]; changed = true; }
                }
                if (flat) {
                    if (valid) {
                        const int pi = (int)(pos & 1023u), pj = (int)((pos >> 10) & 1023u), pk = (int)(pos >> 20);
                        if (changed) {
                            const int lv = (di > 0 ? pi - 1 : ni - 2 - pi) + (dj > 0 ? pj - 1 : nj - 2 - pj) + (dk > 0 ? pk - 1 : nk - 2 - pk);
                            atomicAdd(&s_lvcnt[lv], 1u);
                        } else {
                            const int p = pi + ni * (pj + nj * pk);
                            atomicAnd(&s_act[p >> 5], ~(1u << (p & 31)));
                        }
                    }
                } else if (changed) {
                    K[n] = ((u64)__float_as_uint(phi) << 32) | (sweep_code(s) << kCodeShift) | (uint32_t)ct;
                    ++n_changed;
                    sweep_record_change<SP>(pos, s, di, dj, dk, ni, nj, nk, dstride, s_dirty, chg_code, s_act, s_lvcnt);
                }
            };
            int ln = 0;                                 // SP: nodes waiting in s_list (warp-uniform)
            int fnext = (crank * (NT / 32) + warp) * 32, fidx = 0;            // flat visit: next chunk of bitmap words, this lane's word
            uint32_t fword = 0u;                        //             and what is left of it
            for (int qb = beg + crank * NT + warp * 32; ; qb += R * NT * csize) {
                bool vA = false, vB = false, kp0 = false, kp1 = false;
                uint32_t posA = 0, posB = 0, keep0 = 0, keep1 = 0;
                auto node_at = [&](int q, uint32_t& pos) -> bool {
                    if (q >= end) return false;
                    const uint32_t o = __ldg(order + q);
                    const int a = o & 1023, b = (o >> 10) & 1023, c = o >> 20;
                    const int i = di > 0 ? 1 + a : ni - 2 - a;
                    const int j = dj > 0 ? 1 + b : nj - 2 - b;
                    const int k = dk > 0 ? 1 + c : nk - 2 - c;
                    pos = (uint32_t)i | ((uint32_t)j << 10) | ((uint32_t)k << 20);
                    if (!SP) return true;
                    const int p = i + ni * (j + nj * k);
                    return (((const volatile uint32_t*)s_act)[p >> 5] >> (p & 31)) & 1u;
                };
                if (SP) {
                    // scan R rounds against the activity bits. Where most nodes are active they go straight on; otherwise
                    // the active ones collect in s_list and a batch runs when 32 R are waiting, or with what is left once
                    // the range has been scanned
                    bool more = qb < end;
                    bool direct = false;
                    if (flat) {
                        // the speculative visit takes the marked nodes straight from the bitmap: a warp holds 32 words at a
                        // time, every lane hands over the lowest bit of its word per step (the order of the visit is free)
                        more = true;
                        if (!__any_sync(0xffffffffu, fword != 0u)) {
                            if (fnext < act_words) {
                                fidx = fnext + lane;
                                fword = fidx < act_words ? ((const volatile uint32_t*)s_act)[fidx] : 0u;
                                fnext += NT * csize;
                                continue;
                            }
                            more = false;
                            if (ln == 0) break;
                        } else {
                            vA = fword != 0u;
                            if (vA) {
                                const int p = 32 * fidx + __ffs(fword) - 1;
                                fword &= fword - 1u;
                                // (p < 65,536: the float quotients are exact, as in the band init's box walk)
                                const int r = (int)(((float)p + 0.5f) * rcp_ni), k = (int)(((float)r + 0.5f) * rcp_nj);
                                const int i = p - r * ni, j = r - k * nj;
                                posA = (uint32_t)i | ((uint32_t)j << 10) | ((uint32_t)k << 20);
                                // bits of nodes outside this sweep's lattice (grid faces, tail of the last word) are not nodes to visit
                                vA = k < nk && (di > 0 ? i >= 1 : i <= ni - 2) && (dj > 0 ? j >= 1 : j <= nj - 2) && (dk > 0 ? k >= 1 : k <= nk - 2);
                            }
                            const uint32_t m0 = __ballot_sync(0xffffffffu, vA);
                            if (vA) s_list[ln + __popc(m0 & lt)] = posA;
                            ln += __popc(m0);
                            if (ln < 32 * R) continue;
                        }
                    } else if (more) {
                        vA = node_at(qb + lane, posA);
                        if (R == 2) vB = node_at(qb + NT * csize + lane, posB);
                        const uint32_t m0 = __ballot_sync(0xffffffffu, vA), m1 = R == 2 ? __ballot_sync(0xffffffffu, vB) : 0u;
                        const int c0 = __popc(m0), c1 = __popc(m1);
                        if (c0 + c1 == 0) continue;
                        direct = ln == 0 && c0 + c1 >= 20 * R;
                        if (!direct) {
                            if (vA) s_list[ln + __popc(m0 & lt)] = posA;
                            if (R == 2 && vB) s_list[ln + c0 + __popc(m1 & lt)] = posB;
                            ln += c0 + c1;
                            if (ln < 32 * R) continue;
                        }
                    } else if (ln == 0) {
                        break;
                    }
                    if (!direct) {
                        __syncwarp();
                        const int cn = min(ln, 32 * R);
                        vA = lane < cn;
                        if (vA) posA = s_list[lane];
                        if (R == 2) { vB = 32 + lane < cn; if (vB) posB = s_list[32 + lane]; }
                        // fewer than 32 R nodes stay behind, in registers while the queue is in use
                        kp0 = cn + lane < ln;
                        if (kp0) keep0 = s_list[cn + lane];
                        if (R == 2) { kp1 = cn + 32 + lane < ln; if (kp1) keep1 = s_list[cn + 32 + lane]; }
                        __syncwarp();
                        ln -= cn;
                    }
                } else {
                    if (qb >= end) break;
                    vA = node_at(qb + lane, posA);
                    if (R == 2) vB = node_at(qb + NT * csize + lane, posB);
                }
                uint32_t needA, needB = 0;
                float phiA, phiB = 0.f;
                int ctA, nA, offA, ctB = -1, nB = 0, offB = 0;
                int W = node_phase(vA, posA, 0, needA, phiA, ctA, nA, offA);
                if (R == 2 && (SP ? __any_sync(0xffffffffu, vB) : qb + NT * csize < end)) W += node_phase(vB, posB, W, needB, phiB, ctB, nB, offB);
                if (W != 0) {
                    __syncwarp();
                    for (int e = lane; e < W; e += 32) {
                        uint32_t pp = q_pos[e];
                        V3 gx = node_pos((int)(pp & 1023u), (int)((pp >> 10) & 1023u), (int)(pp >> 20), P);
                        q_pos[e] = __float_as_uint(point_triangle_distance(gx, load_trirec(T + 4 * (int64_t)q_tri[e])));
                        ++n_eval;
                    }
                    __syncwarp();
                }
                if (W != 0 || flat) {
                    apply(vA, needA, posA, phiA, ctA, nA, offA);
                    if (R == 2) apply(vB, needB, posB, phiB, ctB, nB, offB);
                    __syncwarp();                   // the queue is free again
                }
                if (SP && kp0) s_list[lane] = keep0;
                if (SP && R == 2 && kp1) s_list[32 + lane] = keep1;
            }
            csync();
        }
        }
        if (n_changed) atomicAdd(s_nchg, n_changed);
    };

    for (int s = s_begin; s < s_end; ++s) {
        // activity bits + speculation once the previous sweep changed few nodes (spec_div: tuning / test knob, 0 = never)
        csync();
        const int prev = *(volatile int*)s_nchg;
        // (spec_div < 0: always, from sweep 1 on -- tests)
        const bool spec = spec_div < 0 ? s >= 1 : (spec_div > 0 && s >= 8 && (int64_t)prev * spec_div <= (int64_t)n_lattice);
        sweep_threshold_words(s, s_thr[s], s_thrw, s_ib);
        csync();
        if (threadIdx.x == 0 && crank == 0) *s_nchg = 0;
        if (spec) sweep(std::true_type(), s); else sweep(std::false_type(), s);
    }
    if (counter && n_eval) atomicAdd(counter + 1, n_eval);
}

// -------------------------------------------------------------- finalize ----
__device__ __forceinline__ float log_transform(float x)
{
    // losses.py:7-11: sign(x) * log(|x| + 1)
    float s = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
    return s * logf(fabsf(x) + 1.f);
}

__global__ void __launch_bounds__(256)
tdf_finalize_kernel(const u64* __restrict__ key, const uint32_t* __restrict__ parity,
                    const int64_t* __restrict__ vert_off, const int64_t* __restrict__ tri_off, int mesh0, int n_mesh,
                    g3d_tdf_params P, g3d_tdf_outputs O, int signed_mode, int bricked)
{
    const int lane = threadIdx.x & 31;
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // (mesh - mesh0, k, j)
    int64_t nrow = (int64_t)n_mesh * P.nk * P.nj;
    if (row >= nrow) return;
    // (32-bit divisions: a 64-bit one is ~100 instructions, and a thread would make four)
    const uint32_t rows_per_mesh = (uint32_t)P.nk * (uint32_t)P.nj;
    const int mesh = mesh0 + (nrow < 0x7fffffffll ? (int)((uint32_t)row / rows_per_mesh) : (int)(row / rows_per_mesh));
    const uint32_t rin = (uint32_t)(row - (int64_t)(mesh - mesh0) * rows_per_mesh);      // row inside the mesh: j + nj * k
    int64_t nvox = nvox_of(P), pw = (nvox + 31) / 32;
    int64_t rbase = (int64_t)rin * P.ni;                                  // node index of i = 0
    const KeyMap km = keymap_of(P);
    const u64* K = key + (int64_t)mesh * (bricked ? km.stride : nvox);
    const int rk = (int)(rin / (uint32_t)P.nj), rj = (int)(rin - (uint32_t)rk * (uint32_t)P.nj);
    const int kyz = ky(rj, km.sbj) + kz(rk, km.sbk);
    const uint32_t* par = parity + (int64_t)mesh * pw;
    const bool aligned = (P.ni & 31) == 0;
    if (O.status && lane == 0 && rbase == 0 &&
        (tri_off[mesh + 1] == tri_off[mesh] || vert_off[mesh + 1] == vert_off[mesh]))
        atomicOr(O.status + mesh, G3D_MESH_EMPTY);                          // LoaderMesh.h:83
    int carry = 0;
    for (int i0 = 0; i0 < P.ni; i0 += 32) {
        int i = i0 + lane;
        bool valid = i < P.ni;
        int64_t idx = rbase + i;
        uint32_t bit = valid ? ((par[idx >> 5] >> (idx & 31)) & 1u) : 0u;
        uint32_t bal = __ballot_sync(0xffffffffu, bit);
        int incl = __popc(bal & (0xffffffffu >> (31 - lane)));
        bool inside = signed_mode && (((carry + incl) & 1) != 0);           // cpp:176-181
        carry += __popc(bal);
        float sdf = 0.f;
        int ct = -1;
        if (valid) {
            u64 kv = K[bricked ? (int64_t)(kx(i) + kyz) : idx];
            float phi = __uint_as_float((uint32_t)(kv >> 32));
            const uint32_t tmask = bricked ? kTriMask : kTriMaskExact;                 // (bricked = faithful-mode keys)
            ct = (((uint32_t)kv & tmask) == tmask) ? -1 : (int)((uint32_t)kv & tmask);
            if (inside) phi = -phi;
            sdf = mul(phi, P.out_scale);                                    // main.cpp:119
        }
        bool occ = valid && (fabsf(sdf) <= P.occ_thresh);                   // vox2mesh/main.cpp:92
        uint32_t ob = __ballot_sync(0xffffffffu, occ);
        int64_t g = (int64_t)mesh * nvox + idx;
        if (valid) {
            if (O.sdf) O.sdf[g] = sdf;
            float tdf = sdf > P.trunc ? P.trunc : sdf;                      // dataset_loader.py:137-139
            if (O.tdf) O.tdf[g] = tdf;
            if (O.tdflog) O.tdflog[g] = log_transform(tdf);
            if (O.occ_f32) O.occ_f32[g] = occ ? 1.f : 0.f;
            if (O.closest_tri) O.closest_tri[g] = ct;
        }
        if (O.occ_bits) {
            uint32_t* ow = O.occ_bits + (int64_t)mesh * pw;
            if (aligned) { if (lane == 0) ow[idx >> 5] = ob; }
            else if (occ) atomicOr(ow + (idx >> 5), 1u << (idx & 31));
        }
    }
}

__global__ void tdf_from_sdf_kernel(const float* __restrict__ sdf, int64_t n, float trunc,
                                    float* __restrict__ tdf, float* __restrict__ tdflog)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float x = sdf[i];
        x = x > trunc ? trunc : x;
        if (tdf) tdf[i] = x;
        if (tdflog) tdflog[i] = log_transform(x);
    }
}

__global__ void unpack_occ_kernel(const uint32_t* __restrict__ bits, int64_t n_bits, float* __restrict__ out)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_bits; i += stride)
        out[i] = ((bits[i >> 5] >> (i & 31)) & 1u) ? 1.f : 0.f;
}

int grid_for(int64_t n, int block, int max_blocks)
{
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

}  // namespace

namespace {
template <int NT, bool CL>
int launch_sweep_nt(int n_mesh, int csize, int nlev, const TdfWs& w, const int64_t* tri_off, const g3d_tdf_params& p,
                    u64* cnt, cudaStream_t stream, int s_begin = 0, int s_end = 16, uint32_t* chg = nullptr)
{
    const int lev_words = (nlev + 1 + 3) & ~3;
    const size_t smem = (size_t)(lev_words + (NT / 32) * 2 * sweep_rounds(NT) * kSweepQueue + (8 * (nlev + 3) + 3) / 4 +
                                 (p.ni - 1) + (p.nj - 1) + (p.nk - 1)) * 4;
    auto kern = tdf_sweep_kernel<NT, CL>;
    G3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CL && csize > 8) G3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_mesh * csize));
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL ? 1 : 0;
    G3D_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, w.key, (const float4*)w.tri, tri_off, (const uint32_t*)w.order,
                                    (const int32_t*)w.level_start, nlev, lev_words, p, cnt, s_begin, s_end, chg, w.act_words));
    return G3D_OK;
}

// the second-pass kernel (activity bits + speculation): one CTA per mesh with the bits in shared memory, or one cluster
// per mesh with the bits in global memory
template <int NT, bool CL>
int launch_sweep_bm(int n_mesh, int csize, int nlev, const TdfWs& w, const int64_t* tri_off, const g3d_tdf_params& p, u64* cnt,
                    cudaStream_t stream, int s_begin)
{
    const int lev_words = (nlev + 1 + 3) & ~3;
    const size_t smem = (size_t)(lev_words + (NT / 32) * 2 * sweep_rounds(NT) * kSweepQueue + (8 * (nlev + 3) + 3) / 4 +
                                 (CL ? 0 : w.act_words + nlev + 4)) * 4;
    auto kern = tdf_sweep_bm_kernel<NT, CL>;
    G3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CL && csize > 8) G3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    // speculate once a sweep changed at most 1/8 of the lattice nodes; G3D_SWEEP_SPEC overrides the divisor (0 = never)
    const char* e_sp = getenv("G3D_SWEEP_SPEC");
    const int spec_div = e_sp ? atoi(e_sp) : 8;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_mesh * csize));
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL ? 1 : 0;
    G3D_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, w.key, (const float4*)w.tri, tri_off, (const uint32_t*)w.order,
                                    (const int32_t*)w.level_start, nlev, lev_words, p, cnt, w.chg, (const uint32_t*)w.bnd,
                                    w.act_words, s_begin, 16, spec_div, w.gact, w.glvcnt, (int*)w.gnchg));
    return G3D_OK;
}

template <int NT>
int sweep_two_kernels(int n_mesh, int nlev, const TdfWs& w, const int64_t* tri_off, const g3d_tdf_params& p, u64* cnt,
                      cudaStream_t stream, int split)
{
    if (split > 0) {
        int rc = launch_sweep_nt<NT, false>(n_mesh, 1, nlev, w, tri_off, p, cnt, stream, 0, split, w.chg);
        if (rc != G3D_OK) return rc;
    }
    return split < 16 ? launch_sweep_bm<NT, false>(n_mesh, 1, nlev, w, tri_off, p, cnt, stream, split) : G3D_OK;
}

int launch_sweep(int n_mesh, int nlev, const TdfWs& w, const int64_t* tri_off, const g3d_tdf_params& p, u64* cnt,
                 cudaStream_t stream)
{
    // Threads per mesh. Small CTAs win for 32^3-class grids (measured on B200, 1,024 meshes: 128 thr 22.2 ms,
    // 256 thr 24.4, 512 thr 28.4, 1024 thr 31.1): more meshes are resident per SM, so one mesh's barrier
    // wait is covered by another mesh's arithmetic. Large grids get more threads per mesh.
    // test / tuning overrides, read per call (the tests flip them between calls)
    const char* e_nt = getenv("G3D_SWEEP_THREADS");
    const char* e_cl = getenv("G3D_SWEEP_CLUSTER");
    int forced = e_nt ? atoi(e_nt) : 0, forced_cl = e_cl ? atoi(e_cl) : 0;
    if (forced != 128 && forced != 256 && forced != 512 && forced != 1024) forced = 0;
    if (forced_cl != 2 && forced_cl != 4 && forced_cl != 8 && forced_cl != 16) forced_cl = 0;
    int64_t widest = (int64_t)(p.ni - 1) * (p.nj - 1);           // upper bound on a level's node count
    int nt = widest >= 16384 ? 1024 : (widest >= 4096 ? 512 : (widest >= 2048 ? 256 : 128));
    // a batch too small to fill the GPU with small CTAs is latency-bound per mesh: give each mesh more threads,
    // and a whole cluster of CTAs when one CTA cannot cover a level
    const int sms = sm_count();
    int csize = 1;
    if (n_mesh < sms) nt = 1024;
    else if (n_mesh < 2 * sms) nt = nt < 512 ? 512 : nt;
    else if (n_mesh < 4 * sms) nt = nt < 256 ? 256 : nt;
    if (nt == 1024 && widest > 2048) {
        while (csize < 8 && (int64_t)csize * 1024 < widest / 2 && (int64_t)n_mesh * csize * 2 <= sms) csize *= 2;
    }
    // 16 CTAs (a non-portable cluster size, which B200 schedules) once a level holds more nodes than 12 CTAs have threads:
    // one round per thread and level instead of two (128^3, one mesh: 20.5 -> 17.5 ms)
    if (csize == 8 && widest > 12288 && (int64_t)n_mesh * 16 * 2 <= sms) csize = 16;
    if (forced) nt = forced;
    if (forced_cl) { csize = forced_cl; nt = 1024; }
    if (csize == 16) {                              // can the device place such a cluster at all?
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(16); q.blockDim = dim3(1024);
        q.dynamicSmemBytes = (size_t)(((nlev + 1 + 3) & ~3) + 32 * 2 * sweep_rounds(1024) * kSweepQueue + (8 * (nlev + 3) + 3) / 4 +
                                      (p.ni - 1) + (p.nj - 1) + (p.nk - 1)) * 4;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = 16; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        q.attrs = qa; q.numAttrs = 1;
        int n_cl = 0;
        auto kern = tdf_sweep_kernel<1024, true>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q.dynamicSmemBytes);
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (cudaOccupancyMaxActiveClusters(&n_cl, kern, &q) != cudaSuccess || n_cl < 1) { cudaGetLastError(); csize = 8; }
    }
    const char* e_bm = getenv("G3D_SWEEP_BITMAP");
    const bool bm_off = e_bm && atoi(e_bm) == 0;
    if (csize > 1) {
        // a cluster per mesh: first pass in the plain kernel (recording its changes), second pass in the bitmap kernel
        if (bm_off || w.act_words == 0 || n_mesh > kClusterMeshes)
            return launch_sweep_nt<1024, true>(n_mesh, csize, nlev, w, tri_off, p, cnt, stream);
        G3D_CUDA_TRY(cudaMemsetAsync(w.chg, 0, (size_t)n_mesh * 17 * w.act_words * 4, stream));
        int rc = launch_sweep_nt<1024, true>(n_mesh, csize, nlev, w, tri_off, p, cnt, stream, 0, 8, w.chg);
        if (rc != G3D_OK) return rc;
        return launch_sweep_bm<1024, true>(n_mesh, csize, nlev, w, tri_off, p, cnt, stream, 8);
    }
    // grids of at most 65,536 nodes: the kernel with one activity bit per node in shared memory; G3D_SWEEP_BITMAP=0 turns
    // it off (tests run both)
    const bool bm = w.act_words > 0 && act_in_smem(p) && !bm_off;
    // first pass: the plain kernel, which also puts every change on record; second pass: the bitmap kernel. (G3D_SWEEP_SPLIT:
    // the sweep at which it takes over, 0 = from the start -- tests and tuning.)
    const char* e_split = getenv("G3D_SWEEP_SPLIT");
    const int split = e_split ? max(0, min(16, atoi(e_split))) : 8;
    if (bm) G3D_CUDA_TRY(cudaMemsetAsync(w.chg, 0, (size_t)n_mesh * 17 * w.act_words * 4, stream));
#define G3D_SWEEP_CASE(N) (!bm ? launch_sweep_nt<N, false>(n_mesh, 1, nlev, w, tri_off, p, cnt, stream) \
                               : sweep_two_kernels<N>(n_mesh, nlev, w, tri_off, p, cnt, stream, split))
    switch (nt) {
        case 128: return G3D_SWEEP_CASE(128);
        case 512: return G3D_SWEEP_CASE(512);
        case 1024: return G3D_SWEEP_CASE(1024);
        default: return G3D_SWEEP_CASE(256);
    }
#undef G3D_SWEEP_CASE
}
}  // namespace

int tdf_validate(int32_t n_mesh, int64_t total_verts, int64_t total_tris, const g3d_tdf_params& p)
{
    if (n_mesh < 0 || total_verts < 0 || total_tris < 0) { set_error("negative size"); return G3D_ERR_INVALID; }
    if (p.ni < 1 || p.nj < 1 || p.nk < 1 || p.ni > 1024 || p.nj > 1024 || p.nk > 1024) {
        set_error("grid dims must be in [1, 1024], got %d %d %d", p.ni, p.nj, p.nk);
        return G3D_ERR_INVALID;
    }
    if (nvox_of(p) >= (int64_t)INT_MAX) { set_error("grid too large"); return G3D_ERR_INVALID; }
    if (!(p.dx > 0.f)) { set_error("dx must be positive"); return G3D_ERR_INVALID; }
    if (p.exact_band < 0 || p.exact_band > 1024) { set_error("bad exact_band"); return G3D_ERR_INVALID; }
    return G3D_OK;
}

// The work of one batch call in four phases, so that the host-buffer front end can start on the first
// triangles while the rest of the batch is still crossing PCIe:
//   tdf_begin        validate, carve the workspace, initialise keys / parity
//   tdf_stage_tris   per-triangle work for triangles [t0, t1): records, boxes, parity (+ band init, faithful mode)
//   tdf_solve        the whole-batch part: sweeps, or the exact kernel
//   tdf_finalize     the fused epilogue for meshes [m0, m1), so that results can start their way back per chunk
int tdf_begin(int64_t total_verts, int64_t total_tris, int32_t n_mesh, const g3d_tdf_params& p,
              const g3d_tdf_outputs& out, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    int rc = tdf_validate(n_mesh, total_verts, total_tris, p);
    if (rc != G3D_OK) return rc;
    if (p.mode != G3D_MODE_FAITHFUL && p.mode != G3D_MODE_EXACT) { set_error("tdf: unknown mode %d", p.mode); return G3D_ERR_INVALID; }
    // Per MESH the limit is 2^22 - 2 triangles, 2^27 - 2 in exact mode (closest_tri shares a 32-bit word with bookkeeping bits; checked
    // on the device, where the offsets live: G3D_MESH_TOO_MANY_TRIS). Per CALL the evaluation queues hold global
    // triangle numbers in 32 bits.
    if (total_tris >= ((int64_t)1 << 32) - 1) {
        set_error("at most 2^32 - 2 triangles per call, got %lld", (long long)total_tris);
        return G3D_ERR_INVALID;
    }
    if (n_mesh == 0) return G3D_OK;
    const size_t need = p.mode == G3D_MODE_EXACT ? exact::carve(nullptr, n_mesh, total_tris, p).bytes
                                                 : carve(nullptr, n_mesh, total_tris, p).bytes;
    if (need > ws_bytes || (!ws && need)) {
        set_error("workspace too small: need %zu bytes, got %zu", need, ws_bytes);
        return G3D_ERR_WORKSPACE;
    }
    const int sms = sm_count();
    const int64_t nvox = nvox_of(p), pw = (nvox + 31) / 32;
    const float far = (float)(p.ni + p.nj + p.nk) * p.dx;                  // cpp:123
    union { float f; uint32_t u; } fb; fb.f = far;
    const u64 init = ((u64)fb.u << 32) | 0xffffffffull;                      // closest_tri = -1
    prof_begin(stream);
    u64* key; uint32_t* parity;
    if (p.mode == G3D_MODE_EXACT) {
        exact::Ws w = exact::carve(ws, n_mesh, total_tris, p);
        key = w.key; parity = w.parity;
        G3D_CUDA_TRY(cudaMemsetAsync(w.cursor, 0, (size_t)n_mesh * (exact::ncell_of(p) + 1) * 4, stream));
        G3D_CUDA_TRY(cudaMemsetAsync(w.dil, 0, (size_t)n_mesh * 4, stream));
        G3D_CUDA_TRY(cudaMemsetAsync(w.occ, 0, (size_t)n_mesh * ((exact::ncell_of(p) + 31) / 32) * 4, stream));
    } else {
        TdfWs w = carve(ws, n_mesh, total_tris, p);
        key = w.key; parity = w.parity;
    }
    const int64_t n_key = (int64_t)n_mesh * (p.mode == G3D_MODE_EXACT ? nvox : keymap_of(p).stride);
    tdf_fill_kernel<<<grid_for(n_key, 256, sms * 16), 256, 0, stream>>>(
        key, n_key, init, parity, n_mesh * pw, out.status, n_mesh,
        (out.occ_bits && (p.ni & 31)) ? out.occ_bits : nullptr, n_mesh * pw);
    prof_mark(ST_FILL, stream);
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

int tdf_stage_tris(const float* verts, const int64_t* vert_off, const void* tris, int index_bytes, const int64_t* tri_off,
                   int64_t t0, int64_t t1, int32_t m0, int32_t m1, int64_t total_tris, int32_t n_mesh, const g3d_tdf_params& p,
                   const g3d_tdf_outputs& out, void* ws, cudaStream_t stream)
{
    if (index_bytes != 4 && index_bytes != 2) { set_error("tdf: triangle indices must be 2 or 4 bytes wide"); return G3D_ERR_INVALID; }
    if (n_mesh == 0 || t1 <= t0) return G3D_OK;
    const int sms = sm_count();
    const float far = (float)(p.ni + p.nj + p.nk) * p.dx;
    if (p.mode == G3D_MODE_EXACT) {
        exact::Ws w = exact::carve(ws, n_mesh, total_tris, p);
        // 6 CTAs per SM (40 registers, a few spills) beat 3 (76 registers): the kernel waits on its gathers and atomics,
        // measured 2.22 against 2.61 ms per 1,024 meshes
#define G3D_PREP_LAUNCH(IT, MINB) tdf_prep_kernel<IT, MINB><<<(unsigned)((t1 - t0 + 255) / 256), 256, 0, stream>>>( \
            verts, vert_off, (const IT*)tris, tri_off, t0, t1, n_mesh, p, w.tri, nullptr, w.parity, out.status, w.fast, w.tmp, w.cursor, w.dil)
        if (index_bytes == 2) G3D_PREP_LAUNCH(uint16_t, 6);
        else G3D_PREP_LAUNCH(uint32_t, 6);
#undef G3D_PREP_LAUNCH
        prof_mark(ST_PREP, stream);
    } else {
        TdfWs w = carve(ws, n_mesh, total_tris, p);
        // resident CTAs per SM the kernel is compiled for: 4 (64 registers) measured 0.72 ms per 1,024 meshes, 3 (76): 0.78,
        // 6 (40, spills): 0.76 -- the kernel waits on its vertex gathers. G3D_PREP_MINB overrides (tuning aid).
        const char* e_pb = getenv("G3D_PREP_MINB");
        const int pb = e_pb ? atoi(e_pb) : 4;
#define G3D_PREP_F(IT, MB) tdf_prep_kernel<IT, MB><<<(unsigned)((t1 - t0 + 255) / 256), 256, 0, stream>>>( \
            verts, vert_off, (const IT*)tris, tri_off, t0, t1, n_mesh, p, w.tri, w.box, w.parity, out.status, nullptr, nullptr, nullptr, nullptr)
        if (index_bytes == 2) { if (pb >= 6) G3D_PREP_F(uint16_t, 6); else if (pb >= 4) G3D_PREP_F(uint16_t, 4); else G3D_PREP_F(uint16_t, 3); }
        else { if (pb >= 6) G3D_PREP_F(uint32_t, 6); else if (pb >= 4) G3D_PREP_F(uint32_t, 4); else G3D_PREP_F(uint32_t, 3); }
#undef G3D_PREP_F
        prof_mark(ST_PREP, stream);
        // [t0, t1) are the triangles of meshes [m0, m1). Batches of small grids: one CTA per 16^3 block of a mesh, keys in
        // shared memory; each block scans all band boxes of its mesh, so big grids keep the scatter kernel, and so do
        // batches too small to give every SM a block.
        const int nbx = (p.ni + kTile - 1) / kTile, nby = (p.nj + kTile - 1) / kTile, nbz = (p.nk + kTile - 1) / kTile;
        const char* env_tiled = getenv("G3D_INIT_TILED");   // 0/1 overrides the choice (read per call: tests flip it)
        const int tiled = env_tiled ? (atoi(env_tiled) ? 1 : 0) : 2;
        const bool use_tiled = tiled == 2 ? ((int64_t)nbx * nby * nbz <= 64 && (int64_t)(m1 - m0) * nbx * nby * nbz >= 2 * sms) : tiled == 1;
        if (use_tiled && (int64_t)(m1 - m0) * nbx * nby * nbz < (int64_t)INT_MAX)
            tdf_init_tiled_kernel<<<(unsigned)((int64_t)(m1 - m0) * nbx * nby * nbz), 256, 0, stream>>>(
                w.tri, w.box, tri_off, m0, nbx, nby, nbz, p, w.key, far, prof_counters());
        else {
            tdf_init_kernel<<<grid_for((t1 - t0) * 32, 256, sms * 32), 256, 0, stream>>>(
                w.tri, w.box, tri_off, t0, t1, p, w.key, far, prof_counters());
            const int64_t n_keys = (int64_t)(m1 - m0) * nvox_of(p);
            tdf_band_flags_kernel<<<(unsigned)((n_keys + 255) / 256), 256, 0, stream>>>(w.key, w.box, tri_off, m0, m1 - m0, p);
        }
        prof_mark(ST_INIT, stream);
    }
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

int tdf_solve(const int64_t* tri_off, int64_t total_tris, int32_t n_mesh, const g3d_tdf_params& p, void* ws,
              cudaStream_t stream)
{
    if (n_mesh == 0) return G3D_OK;
    u64* cnt = prof_counters();
    if (p.mode == G3D_MODE_EXACT) {
        exact::Ws w = exact::carve(ws, n_mesh, total_tris, p);
        if (total_tris > 0) {
            const int64_t ncell = exact::ncell_of(p), n_groups = (int64_t)n_mesh * ncell;
            const float far = (float)(p.ni + p.nj + p.nk) * p.dx;
            // distances up to R matter: trunc is in output units (phi * out_scale)
            float R = __builtin_inff();
            if (p.trunc < 3.0e38f && p.out_scale > 0.f) R = p.trunc / p.out_scale * 1.0001f;
            // filter margin of the two-stage evaluation, scaled by the extent of the coordinates involved
            float S = 0.f;
            const int dims[3] = { p.ni, p.nj, p.nk };
            for (int a = 0; a < 3; ++a) {
                S = fmaxf(S, fabsf(p.origin[a]));
                S = fmaxf(S, fabsf(p.origin[a] + (float)dims[a] * p.dx));
            }
            const float kappa = 1e-10f * S * S;
            // the cell index: CSR starts, entries in cell order, cell boxes
            exact::exact_scan_kernel<<<(unsigned)n_mesh, 1024, 0, stream>>>(w.cursor, w.cell_start, ncell);
            exact::exact_fill_kernel<<<(unsigned)((total_tris + 255) / 256), 256, 0, stream>>>(
                w.tmp, tri_off, total_tris, n_mesh, ncell, w.cell_start, w.cursor, w.ent);
            exact::exact_cellbox_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, stream>>>(
                w.ent, tri_off, w.cell_start, ncell, n_groups, w.cellbox, w.occ);
            prof_mark(ST_ORDER, stream);
            // voxel group per warp: 1x1x1 cells (8 voxels), 2x1x1 (16) or 2x2x1 (32); G3D_EXACT_GROUP overrides (tuning)
            // the largest group that still gives the GPU two waves of warps (a single 32^3 mesh has only 1,024 groups of 32 voxels)
            const char* e_grp = getenv("G3D_EXACT_GROUP");
            const int ncx = exact::cells_of(p.ni), ncy = exact::cells_of(p.nj), ncz = exact::cells_of(p.nk);
            const int64_t enough = (int64_t)sm_count() * 64;
            int grp = 32;
            if ((int64_t)n_mesh * ((ncx + 1) / 2) * ((ncy + 1) / 2) * ncz < enough) grp = 16;
            if (grp == 16 && (int64_t)n_mesh * ((ncx + 1) / 2) * ncy * ncz < enough) grp = 8;
            if (e_grp) grp = atoi(e_grp);
            const int gcx = grp >= 16 ? 2 : 1, gcy = grp >= 32 ? 2 : 1;
            const int64_t n_wgroups = (int64_t)n_mesh * ((ncx + gcx - 1) / gcx) * ((ncy + gcy - 1) / gcy) * ncz;
            const int64_t want = (n_wgroups + exact::kExactWarps - 1) / exact::kExactWarps;
            const int64_t cap = (int64_t)sm_count() * 64;        // persistent warps: 16 CTAs per SM, 4 rounds of slack
            const char* e_occ = getenv("G3D_EXACT_OCC");        // tuning override: resident CTAs per SM the kernel is compiled for
            const int occv = e_occ ? atoi(e_occ) : 8;
            const unsigned nb = (unsigned)(want < cap ? want : cap);
#define G3D_EXACT_LAUNCH(MINB, CX, CY) exact::tdf_exact3_kernel<MINB, CX, CY, 1><<<nb, exact::kExactWarps * 32, 0, stream>>>( \
                w.tri, w.fast, w.ent, w.cell_start, w.cellbox, w.occ, w.dil, tri_off, w.parity, w.key, p, R, far, kappa, n_wgroups, cnt)
            // measured on B200, 1,024 meshes of 19,440 triangles, kernel time: 8 voxels per warp 21.7 ms, 16: 18.4, 32: 17.2 (at 8 CTAs
            // per SM = 64 registers; 7: 17.7, 6: 18.7, 5: 19.3)
            if (grp >= 32) { if (occv >= 8) G3D_EXACT_LAUNCH(8, 2, 2); else G3D_EXACT_LAUNCH(5, 2, 2); }
            else if (grp >= 16) { if (occv >= 8) G3D_EXACT_LAUNCH(8, 2, 1); else G3D_EXACT_LAUNCH(5, 2, 1); }
            else { if (occv >= 8) G3D_EXACT_LAUNCH(8, 1, 1); else G3D_EXACT_LAUNCH(5, 1, 1); }
#undef G3D_EXACT_LAUNCH
            prof_mark(ST_EXACT, stream);
        }
    } else {
        TdfWs w = carve(ws, n_mesh, total_tris, p);
        const int nlev = nlev_of(p);
        if (total_tris > 0 && nlev > 0) {
            size_t smem = (size_t)(nlev + 1 + (p.ni - 1) + (p.nj - 1) - 1) * 4;
            const int64_t n_lat = (int64_t)(p.ni - 1) * (p.nj - 1) * (p.nk - 1);
            const unsigned ob = (unsigned)std::min<int64_t>(std::max<int64_t>(n_lat / 16384, 1), 64);     // 16 entries per thread and more
            tdf_order_kernel<<<ob, 1024, smem, stream>>>(p.ni - 1, p.nj - 1, p.nk - 1, nlev, w.order, w.level_start, w.bnd, w.act_words);
            prof_mark(ST_ORDER, stream);
            int rc = launch_sweep(n_mesh, nlev, w, tri_off, p, cnt, stream);
            if (rc != G3D_OK) return rc;
            prof_mark(ST_SWEEP, stream);
        }
    }
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

// fused epilogue for meshes [m0, m1): signs, scaling, occupancy, truncation, log transform, all outputs
int tdf_finalize(const int64_t* vert_off, const int64_t* tri_off, int64_t total_tris, int32_t n_mesh, int32_t m0,
                 int32_t m1, const g3d_tdf_params& p, const g3d_tdf_outputs& out, void* ws, cudaStream_t stream)
{
    if (n_mesh == 0 || m1 <= m0) return G3D_OK;
    const u64* key; const uint32_t* parity;
    if (p.mode == G3D_MODE_EXACT) { exact::Ws w = exact::carve(ws, n_mesh, total_tris, p); key = w.key; parity = w.parity; }
    else { TdfWs w = carve(ws, n_mesh, total_tris, p); key = w.key; parity = w.parity; }
    tdf_finalize_kernel<<<(unsigned)(((int64_t)(m1 - m0) * p.nk * p.nj * 32 + 255) / 256), 256, 0, stream>>>(
        key, parity, vert_off, tri_off, m0, m1 - m0, p, out, 1, p.mode == G3D_MODE_EXACT ? 0 : 1);
    prof_mark(ST_FINAL, stream);
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

size_t tdf_workspace_bytes(int32_t n_mesh, int64_t total_tris, const g3d_tdf_params& p)
{
    return p.mode == G3D_MODE_EXACT ? exact::carve(nullptr, n_mesh, total_tris, p).bytes : carve(nullptr, n_mesh, total_tris, p).bytes;
}

int tdf_from_sdf_launch(const float* sdf, int64_t n, float trunc, float* tdf, float* tdflog, cudaStream_t stream)
{
    if (n < 0 || (!sdf && n)) { set_error("bad sdf/n"); return G3D_ERR_INVALID; }
    if (n == 0) return G3D_OK;
    tdf_from_sdf_kernel<<<grid_for(n, 256, sm_count() * 16), 256, 0, stream>>>(sdf, n, trunc, tdf, tdflog);
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

int unpack_occ_launch(const uint32_t* bits, int64_t n_bits, float* outp, cudaStream_t stream)
{
    if (n_bits < 0 || ((!bits || !outp) && n_bits)) { set_error("bad bits/out"); return G3D_ERR_INVALID; }
    if (n_bits == 0) return G3D_OK;
    unpack_occ_kernel<<<grid_for(n_bits, 256, sm_count() * 16), 256, 0, stream>>>(bits, n_bits, outp);
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

}  // namespace g3d

// g3d_tdf_exact.cuh -- "exact" mode: brute-force nearest triangle with truncation culling.
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
// One CTA owns an 8x8x8 block of voxels of one mesh (16 warps x a 4x4x2 brick, one voxel per lane)
// and streams the mesh's triangle bounding boxes through shared memory in tiles, double-buffered
// with 1-D TMA bulk copies (cp.async.bulk + mbarrier): while the warps work on tile i, tile i+1 is
// in flight.
//   phase A  all threads: tile boxes vs the block's box with the block's static radius -> compacted
//            survivor list in shared memory;
//   phase B  each warp: survivors vs its brick's box with the brick's RUNNING bound (the largest
//            current best distance among its 32 voxels, a warp-shuffle max-reduce), 32 at a time
//            via ballot; every survivor is then evaluated by all 32 lanes (the triangle record is a
//            broadcast load) with the reference's distance function and merged with
//            (d, t) < (best, best_t) lexicographic order = "lowest triangle index wins ties".
#pragma once

namespace exact {

// block shape in warps (each warp = a 4x4x2 brick) and triangle-box tile size. Small blocks and small tiles win:
// a block barrier costs the slowest brick's time (a brick next to the surface rates thousands of pairs, a far one
// none), and occupancy is bounded by the tile stages in shared memory. Measured on B200, 1,024 meshes, two-stage
// kernel: 2x2x4 warps / 1024 boxes 105 ms; 2x1x2 / 512 (two barriers per tile) 67.7; one barrier per tile with
// three stages: 2x1x2 / 512 90.4, / 256 56.4, / 128 67.1; 2x2x2 / 256 58.9; prefetching the next survivor's
// record costs more than it hides (60.2 vs 56.4).
constexpr int kWX = 2, kWY = 1, kWZ = 2;
constexpr int kTile = 256;
constexpr bool kPrefetchNext = false;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// squared distance between two axis-aligned boxes (0 when they overlap); FMA is fine here, the
// result only feeds a conservative comparison
__device__ __forceinline__ float box_gap2(float4 alo, float4 ahi, V3 blo, V3 bhi)
{
    float gx = fmaxf(0.f, fmaxf(alo.x - bhi.x, blo.x - ahi.x));
    float gy = fmaxf(0.f, fmaxf(alo.y - bhi.y, blo.y - ahi.y));
    float gz = fmaxf(0.f, fmaxf(alo.z - bhi.z, blo.z - ahi.z));
    return fmaf(gx, gx, fmaf(gy, gy, gz * gz));
}

// neutralised triangles carry an inverted infinite box (gap = inf): culled unless the bound is infinite,
// in which case their NaN distances simply never win
__device__ __forceinline__ float bound2_of(float r) { return fmaf(r * r, 1.001f, 1e-12f); }

template <int WX, int WY, int WZ, int TILE>
__global__ void __launch_bounds__(WX * WY * WZ * 32)
tdf_exact_kernel(const float4* __restrict__ tri, const float4* __restrict__ aabb,
                 const int64_t* __restrict__ tri_off, const uint32_t* __restrict__ parity, u64* __restrict__ key,
                 g3d_tdf_params P, float R, float far, int nbx, int nby, int nbz, u64* counter)
{
    constexpr int NT = WX * WY * WZ * 32;
    constexpr int BX = WX * 4, BY = WY * 4, BZ = WZ * 2;                      // nodes per block and axis
    extern __shared__ __align__(128) unsigned char s_raw[];
    float4* s_box = reinterpret_cast<float4*>(s_raw);                         // [2][TILE * 2]
    uint32_t* s_list = reinterpret_cast<uint32_t*>(s_raw + 2 * TILE * 32);    // [TILE]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_list + TILE);             // [2]
    int* s_count = reinterpret_cast<int*>(s_bar + 2);                         // [2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nblk = nbx * nby * nbz;
    const int mesh = blockIdx.x / nblk;
    int b = blockIdx.x - mesh * nblk;
    const int cbx = b % nbx; b /= nbx;
    const int cby = b % nby, cbz = b / nby;
    const int ni = P.ni, nj = P.nj, nk = P.nk;
    const int64_t nvox = nvox_of(P);

    // block = BX x BY x BZ nodes; warp -> 4x4x2 brick; lane -> node
    const int bi0 = cbx * BX + (warp % WX) * 4, bj0 = cby * BY + ((warp / WX) % WY) * 4, bk0 = cbz * BZ + (warp / (WX * WY)) * 2;
    const int i = bi0 + (lane & 3), j = bj0 + ((lane >> 2) & 3), k = bk0 + (lane >> 4);
    const bool valid = i < ni && j < nj && k < nk;
    bool inside = false;
    if (valid) {                                    // running parity along i (makelevelset3.cpp:174-182)
        const uint32_t* par = parity + (int64_t)mesh * ((nvox + 31) / 32);
        int64_t r0 = (int64_t)ni * (j + (int64_t)nj * k), r1 = r0 + i;
        int cnt = 0;
        for (int64_t w = r0 >> 5; w <= (r1 >> 5); ++w) {
            uint32_t m = par[w];
            if (w == (r0 >> 5)) m &= 0xffffffffu << (r0 & 31);
            if (w == (r1 >> 5)) m &= 0xffffffffu >> (31 - (r1 & 31));
            cnt += __popc(m);
        }
        inside = cnt & 1;
    }
    const V3 gx = node_pos(i, j, k, P);
    // like the reference's phi, the running minimum starts at `far` (makelevelset3.cpp:123): a triangle farther
    // away than that never registers. Outside voxels may stop at the truncation radius instead.
    float best = valid ? (inside ? far : fminf(R, far)) : 0.f;
    uint32_t best_t = 0xffffffffu;

    // boxes of the brick and of the block (node positions are monotone in the index)
    const V3 brick_lo = node_pos(bi0, bj0, bk0, P);
    const V3 brick_hi = node_pos(min(bi0 + 3, ni - 1), min(bj0 + 3, nj - 1), min(bk0 + 1, nk - 1), P);
    const V3 blk_lo = node_pos(cbx * BX, cby * BY, cbz * BZ, P);
    const V3 blk_hi = node_pos(min(cbx * BX + BX - 1, ni - 1), min(cby * BY + BY - 1, nj - 1), min(cbz * BZ + BZ - 1, nk - 1), P);
    const int any_inside = __syncthreads_or(inside ? 1 : 0);
    const float blk_bound2 = bound2_of(any_inside ? far : fminf(R, far));

    const int64_t t0 = tri_off[mesh];
    const int T = (int)(tri_off[mesh + 1] - t0);
    const float4* Tm = tri + 4 * t0;
    const float4* Bm = aabb + 2 * t0;
    const int ntile = (T + TILE - 1) / TILE;
    u64 n_pairs = 0, n_tests = 0;

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        s_count[0] = s_count[1] = 0;
    }
    __syncthreads();
    if (tid == 0 && ntile > 0) {
        uint32_t bytes = (uint32_t)min(TILE, T) * 32u;
        mbar_expect_tx(&s_bar[0], bytes);
        bulk_g2s(s_box, Bm, bytes, &s_bar[0]);
    }
    for (int it = 0; it < ntile; ++it) {
        const int st = it & 1;
        const int tile_base = it * TILE;
        const int tn = min(TILE, T - tile_base);
        if (tid == 0 && it + 1 < ntile) {           // stage st^1 was released by the barrier ending iteration it-1
            uint32_t bytes = (uint32_t)min(TILE, T - (tile_base + TILE)) * 32u;
            mbar_expect_tx(&s_bar[st ^ 1], bytes);
            bulk_g2s(s_box + (st ^ 1) * TILE * 2, Bm + 2 * (int64_t)(tile_base + TILE), bytes, &s_bar[st ^ 1]);
        }
        mbar_wait(&s_bar[st], (it >> 1) & 1);
        const float4* box = s_box + st * TILE * 2;

        // ---- phase A: block-level cull, compacted with one shared atomic per warp
        for (int q0 = warp * 32; q0 < tn; q0 += NT) {
            int q = q0 + lane;
            bool pass = false;
            if (q < tn) {
                pass = box_gap2(box[2 * q], box[2 * q + 1], blk_lo, blk_hi) <= blk_bound2;
                ++n_tests;
            }
            uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_count[st], __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pass) s_list[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)q;
            }
        }
        __syncthreads();
        const int nlist = s_count[st];
        if (tid == 0) s_count[st ^ 1] = 0;          // nobody touches the other counter until the next iteration

        // ---- phase B: brick-level cull with the running bound, then the distance evaluations
        for (int l0 = 0; l0 < nlist; l0 += 32) {
            float bound = best;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, d));
            const float bound2 = bound2_of(bound);
            uint32_t tq = 0;                         // the triangle's own index rides in lo.w of its box
            bool pass = false;
            if (l0 + lane < nlist) {
                const int q = (int)s_list[l0 + lane];
                const float4 lo = box[2 * q];
                tq = __float_as_uint(lo.w);
                pass = box_gap2(lo, box[2 * q + 1], brick_lo, brick_hi) <= bound2;
                ++n_tests;
            }
            uint32_t m = __ballot_sync(0xffffffffu, pass);
            while (m) {
                int src = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t t = __shfl_sync(0xffffffffu, tq, src);
                const float d = point_triangle_distance(gx, load_trirec(Tm + 4 * (int64_t)t));
                if (d < best || (d == best && best_t != 0xffffffffu && t < best_t)) { best = d; best_t = t; }
                n_pairs += valid ? 1 : 0;
            }
        }
        __syncthreads();                            // stage st and the list are free again
    }
    if (valid) {
        float phi = (best_t == 0xffffffffu) ? far : best;
        key[(int64_t)mesh * nvox + (i + (int64_t)ni * (j + (int64_t)nj * k))] = ((u64)__float_as_uint(phi) << 32) | best_t;
    }
    if (counter) {
        if (n_pairs) atomicAdd(counter + 2, n_pairs);
        if (n_tests) atomicAdd(counter + 3, n_tests);
    }
}

// ---------------------------------------------------------------------------------------------
// v2: two-stage evaluation. Stage 1 rates every surviving (voxel, triangle) pair with a cheap
// squared distance that is free to use FMA (55 arithmetic instructions, no division, no square
// root: the three clamped edge parameters use precomputed reciprocals and the saturate modifier).
// Only pairs whose cheap distance is within a generous margin of the voxel's running cheap minimum
// can be the minimum of the reference's float function; those few are queued and re-evaluated 32
// at a time with the bit-faithful function (stage 2), which alone decides the result.
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
// always go to stage 2.
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

template <int WX, int WY, int WZ, int TILE>
__global__ void __launch_bounds__(WX * WY * WZ * 32)
tdf_exact2_kernel(const float4* __restrict__ tri, const float4* __restrict__ fast, const float4* __restrict__ aabb,
                  const int64_t* __restrict__ tri_off, const uint32_t* __restrict__ parity, u64* __restrict__ key,
                  g3d_tdf_params P, float R, float far, float eta, float kappa, int nbx, int nby, int nbz, u64* counter)
{
    constexpr int NT = WX * WY * WZ * 32, NW = NT / 32;
    constexpr int BX = WX * 4, BY = WY * 4, BZ = WZ * 2;                      // nodes per block and axis
    extern __shared__ __align__(128) unsigned char s_raw[];
    float4* s_box = reinterpret_cast<float4*>(s_raw);                         // [3][TILE * 2]  TMA stages
    uint32_t* s_list = reinterpret_cast<uint32_t*>(s_raw + 3 * TILE * 32);    // [2][TILE]      block-level survivors
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_list + 2 * TILE);         // [3]
    int* s_count = reinterpret_cast<int*>(s_bar + 3);                         // [3] (+1 pad)
    float* s_wb = reinterpret_cast<float*>(s_count + 4);                      // [NW] running bound^2 of each brick
    uint32_t* s_queue = reinterpret_cast<uint32_t*>(s_wb + NW);               // [NW][64]  t << 5 | owner lane

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t* q = s_queue + warp * 64;
    const int nblk = nbx * nby * nbz;
    const int mesh = blockIdx.x / nblk;
    int b = blockIdx.x - mesh * nblk;
    const int cbx = b % nbx; b /= nbx;
    const int cby = b % nby, cbz = b / nby;
    const int ni = P.ni, nj = P.nj, nk = P.nk;
    const int64_t nvox = nvox_of(P);

    const int bi0 = cbx * BX + (warp % WX) * 4, bj0 = cby * BY + ((warp / WX) % WY) * 4, bk0 = cbz * BZ + (warp / (WX * WY)) * 2;
    const int i = bi0 + (lane & 3), j = bj0 + ((lane >> 2) & 3), k = bk0 + (lane >> 4);
    const bool valid = i < ni && j < nj && k < nk;
    bool inside = false;
    if (valid) {                                    // running parity along i (makelevelset3.cpp:174-182)
        const uint32_t* par = parity + (int64_t)mesh * ((nvox + 31) / 32);
        int64_t r0 = (int64_t)ni * (j + (int64_t)nj * k), r1 = r0 + i;
        int cnt = 0;
        for (int64_t w = r0 >> 5; w <= (r1 >> 5); ++w) {
            uint32_t m = par[w];
            if (w == (r0 >> 5)) m &= 0xffffffffu << (r0 & 31);
            if (w == (r1 >> 5)) m &= 0xffffffffu >> (31 - (r1 & 31));
            cnt += __popc(m);
        }
        inside = cnt & 1;
    }
    const float INF = __int_as_float(0x7f800000);
    const V3 gx = node_pos(i, j, k, P);
    // stage-2 state (decides the result) and stage-1 state (decides who gets a stage-2 evaluation)
    // like the reference's phi, the running minimum starts at `far` (makelevelset3.cpp:123): a triangle farther
    // away than that never registers. Outside voxels may stop at the truncation radius instead.
    const float start = inside ? far : fminf(R, far);
    float best = valid ? start : 0.f;
    uint32_t best_t = 0xffffffffu;
    // pass to stage 2 iff a2 - e2 <= lim2 = 1.002 * ub2 + kappa, ub2 = smallest a2 + e2 seen (start^2 at first)
    float ub2 = valid ? start * start : -1.f;
    float lim2 = valid ? fmaf(ub2, 1.002f, kappa) : -1.f;
    (void)eta; (void)INF;

    const V3 brick_lo = node_pos(bi0, bj0, bk0, P);
    const V3 brick_hi = node_pos(min(bi0 + 3, ni - 1), min(bj0 + 3, nj - 1), min(bk0 + 1, nk - 1), P);
    const V3 blk_lo = node_pos(cbx * BX, cby * BY, cbz * BZ, P);
    const V3 blk_hi = node_pos(min(cbx * BX + BX - 1, ni - 1), min(cby * BY + BY - 1, nj - 1), min(cbz * BZ + BZ - 1, nk - 1), P);

    auto warp_bound2 = [&]() {                      // largest distance^2 that still matters to any voxel of the brick
        float v = lim2;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
        return v * 1.001f;
    };

    const int64_t t0 = tri_off[mesh];
    const int T = (int)(tri_off[mesh + 1] - t0);
    const float4* Tm = tri + 4 * t0;
    const float4* Fm = fast + 5 * t0;
    const float4* Bm = aabb + 2 * t0;
    const int ntile = (T + TILE - 1) / TILE;
    u64 n_pairs = 0, n_tests = 0, n_faithful = 0;
    int qn = 0;

    // stage 2 on 32 queue entries starting at `from` (entries beyond `cnt` are idle lanes)
    auto flush = [&](int from, int cnt) {
        __syncwarp();
        uint32_t ent = 0;
        float d = INF;
        const bool have = lane < cnt;
        if (have) {
            ent = q[from + lane];
            const int o = ent & 31u;
            const V3 go = node_pos(bi0 + (o & 3), bj0 + ((o >> 2) & 3), bk0 + (o >> 4), P);
            d = point_triangle_distance(go, load_trirec(Tm + 4 * (int64_t)(ent >> 5)));
            ++n_faithful;
        }
        for (int r = 0; r < cnt; ++r) {
            const uint32_t er = __shfl_sync(0xffffffffu, ent, r);
            const float dr = __shfl_sync(0xffffffffu, d, r);
            const uint32_t tr = er >> 5;
            if ((int)(er & 31u) == lane && (dr < best || (dr == best && best_t != 0xffffffffu && tr < best_t))) { best = dr; best_t = tr; }
        }
        __syncwarp();
    };
    auto issue_tile = [&](int tile) {               // thread 0: TMA bulk copy of one tile of boxes into stage tile % 3
        const int sg = tile % 3;
        const uint32_t bytes = (uint32_t)min(TILE, T - tile * TILE) * 32u;
        mbar_expect_tx(&s_bar[sg], bytes);
        bulk_g2s(s_box + sg * TILE * 2, Bm + 2 * (int64_t)tile * TILE, bytes, &s_bar[sg]);
    };

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init(&s_bar[2], 1);
        mbar_fence_init();
        s_count[0] = s_count[1] = s_count[2] = 0;
    }
    { const float wb = warp_bound2(); if (lane == 0) s_wb[warp] = wb; }
    __syncthreads();
    if (tid == 0) {
        if (ntile > 0) issue_tile(0);
        if (ntile > 1) issue_tile(1);
    }
    // One block barrier per tile: [phase A of tile it] barrier [phase B of tile it], with three box stages, two
    // survivor lists and three counters so that a fast warp may already run phase A of the next tile.
    for (int it = 0; it < ntile; ++it) {
        const int sg = it % 3, lb = it & 1;
        const int tn = min(TILE, T - it * TILE);
        mbar_wait(&s_bar[sg], (it / 3) & 1);
        const float4* box = s_box + sg * TILE * 2;
        uint32_t* list = s_list + lb * TILE;

        // ---- phase A: block-level cull against the largest running bound of the block's bricks
        float blk_bound2 = s_wb[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) blk_bound2 = fmaxf(blk_bound2, s_wb[w]);
        for (int q0 = warp * 32; q0 < tn; q0 += NT) {
            int qq = q0 + lane;
            bool pass = false;
            if (qq < tn) {
                pass = box_gap2(box[2 * qq], box[2 * qq + 1], blk_lo, blk_hi) <= blk_bound2;
                ++n_tests;
            }
            uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_count[sg], __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pass) list[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)qq;
            }
        }
        if (tid == 0) s_count[(it + 1) % 3] = 0;    // last read in phase B of tile it-2, next written after the barrier
        __syncthreads();
        if (tid == 0 && it + 2 < ntile) issue_tile(it + 2);   // its stage was last read in tile it-1
        const int nlist = s_count[sg];

        // ---- phase B: brick-level cull with the running bound, stage 1 on the survivors
        for (int l0 = 0; l0 < nlist; l0 += 32) {
            const float bound2 = warp_bound2();
            uint32_t tq = 0;                         // the triangle's own index rides in lo.w of its box
            bool pass = false;
            if (l0 + lane < nlist) {
                const int qq = (int)list[l0 + lane];
                const float4 lo = box[2 * qq];
                tq = __float_as_uint(lo.w);
                pass = box_gap2(lo, box[2 * qq + 1], brick_lo, brick_hi) <= bound2;
                ++n_tests;
            }
            uint32_t m = __ballot_sync(0xffffffffu, pass);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t t = __shfl_sync(0xffffffffu, tq, src);
                if (kPrefetchNext && m) {           // pull the next survivor's record while this one is rated
                    const uint32_t tn2 = __shfl_sync(0xffffffffu, tq, __ffs(m) - 1);
                    if (lane < 2) prefetch_l1(reinterpret_cast<const char*>(Fm + 5 * (int64_t)tn2) + lane * 64);
                }
                bool flagged;
                float e2;
                const float a2 = fast_dist2(gx, Fm + 5 * (int64_t)t, flagged, e2);
                n_pairs += valid ? 1 : 0;
                const bool cand = valid && (flagged || !(a2 - e2 > lim2));
                if (!flagged && a2 + e2 < ub2) {     // new upper bound of the minimum: tighten the filter
                    ub2 = a2 + e2;
                    lim2 = fmaf(ub2, 1.002f, kappa);
                }
                const uint32_t cm = __ballot_sync(0xffffffffu, cand);
                if (cm) {
                    if (cand) q[qn + __popc(cm & ((1u << lane) - 1u))] = (t << 5) | (uint32_t)lane;
                    qn += __popc(cm);
                    if (qn >= 32) { qn -= 32; flush(qn, 32); }
                }
            }
        }
        { const float wb = warp_bound2(); if (lane == 0) s_wb[warp] = wb; }
    }
    if (qn > 0) flush(0, qn);
    if (valid) {
        float phi = (best_t == 0xffffffffu) ? far : best;
        key[(int64_t)mesh * nvox + (i + (int64_t)ni * (j + (int64_t)nj * k))] = ((u64)__float_as_uint(phi) << 32) | best_t;
    }
    if (counter) {
        if (n_faithful) atomicAdd(counter + 0, n_faithful);
        if (n_pairs) atomicAdd(counter + 2, n_pairs);
        if (n_tests) atomicAdd(counter + 3, n_tests);
    }
}

struct Ws {
    float4* tri; float4* fast; float4* aabb; int4* box; u64* key; uint32_t* parity; size_t bytes;
};

inline Ws carve(void* base, int32_t n_mesh, int64_t total_tris, const g3d_tdf_params& p)
{
    Ws w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (char*)base + o; };
    int64_t nvox = nvox_of(p), pw = (nvox + 31) / 32;
    w.tri = (float4*)take((size_t)total_tris * 64);
    w.fast = (float4*)take((size_t)total_tris * 80);
    w.aabb = (float4*)take((size_t)total_tris * 32);
    w.box = (int4*)take((size_t)total_tris * 16);
    w.key = (u64*)take((size_t)n_mesh * nvox * 8);
    w.parity = (uint32_t*)take((size_t)n_mesh * pw * 4);
    w.bytes = off;
    return w;
}

}  // namespace exact

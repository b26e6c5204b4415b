// g3d_raycast.cu -- depth view + pose -> 32^3 input occupancy ("visual cone") on sm_100a.
//
// Reproduces compute_index_mapping + raycast_depth (src/scripts/raytracing.py:
// 43-82, 122-144) bit for bit. The reference projects the 8 corners of every
// voxel (262,144 projections per view), takes the clamped pixel bounding box and
// then runs a 32,768-iteration Python loop with a device->host sync per voxel.
// Here one CTA handles (view, z-slab):
//
//  1. the u8 depth raster is read ONCE with 16-byte loads and reduced to a
//     `depth > 0` bit mask in shared memory (H x W/32 words; 8 KB at 256^2),
//     plus two coarser masks holding the OR of every 4 and every 16 consecutive rows;
//  2. the corners are shared lattice points, so each of the (dim+1)^3 lattice
//     points is projected once (7.3x fewer projections) with the reference's
//     arithmetic: cam = pose @ (g2w @ p) accumulated k = 0..3 with one FMA per
//     term (the association MKL sgemm uses for raytracing.py:64, pinned by the
//     golden fixtures), y flipped, u = x*f/|z| + cx, round-half-even, clamp.
//     The first two terms of that sum do not depend on the plane and are formed
//     once per view (3 floats per (x, y) lattice point in shared memory).
//     Planes are projected one at a time and kept two at a time; per plane the
//     4-corner min/max of every (x, y) cell is formed once with packed 16-bit
//     SIMD min/max by the thread that owns the cell in every slice, which keeps
//     the previous plane's value for itself: a voxel's box is that value, four
//     lattice loads and eight packed min/max, and a slice needs one barrier;
//  3. each voxel tests its box against the coarsest fitting mask in at most 4 reads (half-open,
//     clipped to the image as numpy slicing does); a warp ballots 32 voxels
//     into one output word, so the bit-packed grid is written without atomics.
//
// Roofline: HBM. Algorithmic bytes per view = H*W (depth, read once) + 64
// (pose) + dim^3/8 (packed occupancy) = 69,696 B at 256^2, 266,304 B at 512^2.
#include "g3d_internal.h"
#include "g3d_project.cuh"

namespace g3d {

namespace {

__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t w)
{
    // one bit per non-zero byte of w, in byte order
    uint32_t m = __vcmpne4(w, 0u) & 0x08040201u;
    return (m | (m >> 8) | (m >> 16) | (m >> 24)) & 0xfu;
}

// DIM > 0: grid size known at compile time (the reference only ever uses 32): the shared-memory layout and the
// index arithmetic of the inner loops fold into immediates. DIM = 0: any size, from the `dim_arg` argument.
template <int NT, int DIM>
__global__ void __launch_bounds__(NT, 3)
raycast_kernel(const uint8_t* __restrict__ depth, const float* __restrict__ pose, int H, int W,
               float focal, float cx, float cy, int clamp_max, int dim_arg, int zs, float inv_dim,
               uint32_t* __restrict__ occ_bits, float* __restrict__ occ_f32,
               long long* __restrict__ index_map)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int view = blockIdx.x;
    const int z0 = blockIdx.y * zs;
    const int z1 = min(DIM ? DIM : dim_arg, z0 + zs);
    const int Wp = (W + 31) >> 5;
    const int dim = DIM ? DIM : dim_arg;
    const int n = dim + 1, nn = n * n, d2 = dim * dim;
    // fixed-size arrays first, so that their offsets do not depend on the image size
    uint32_t* lat = smem;                           // [2][n*n]  projected lattice plane: u | v << 16
    uint2* quad = reinterpret_cast<uint2*>(lat + 2 * nn + ((2 * nn) & 1));   // [dim*dim] (min2, max2) of a cell's 4 corners in the previous plane
    float* axy = reinterpret_cast<float*>(quad + d2);                        // [3][n*n] the (x, y) part of cam = pose @ g2w @ p
    uint32_t* m1 = reinterpret_cast<uint32_t*>(axy + 3 * nn + ((3 * nn) & 1));   // [H * Wp]  depth > 0, one bit per pixel
    uint32_t* m4 = m1 + (size_t)H * Wp;             // [H * Wp]  OR of rows r .. r+3
    uint32_t* m16 = m4 + (size_t)H * Wp;            // [H * Wp]  OR of rows r .. r+15
    const uint8_t* D = depth + (size_t)view * H * W;
    const int tid = threadIdx.x, lane = tid & 31;

    // ---- 1. depth > 0 bit mask (depth == NULL: index map only)
    if (!depth) {
    } else if ((W & 31) == 0 && ((reinterpret_cast<uintptr_t>(D) & 15) == 0)) {
        const uint4* D4 = reinterpret_cast<const uint4*>(D);
        uint16_t* m16 = reinterpret_cast<uint16_t*>(m1);
        const int nvec = (H * W) >> 4;              // rows are whole words: vec v -> half-word v
        for (int v = tid; v < nvec; v += NT) {
            uint4 x = __ldg(D4 + v);
            uint32_t b = nonzero_nibble(x.x) | (nonzero_nibble(x.y) << 4) | (nonzero_nibble(x.z) << 8) |
                         (nonzero_nibble(x.w) << 12);
            m16[v] = (uint16_t)b;
        }
    } else {
        for (int w = tid; w < H * Wp; w += NT) {
            int row = w / Wp, c0 = (w - row * Wp) << 5;
            uint32_t b = 0;
            for (int c = 0; c < 32 && c0 + c < W; ++c) b |= (D[(size_t)row * W + c0 + c] != 0 ? 1u : 0u) << c;
            m1[w] = b;
        }
    }
    __syncthreads();
    if (depth)
        for (int w = tid; w < H * Wp; w += NT) {
            int row = w / Wp;
            uint32_t b = m1[w];
            if (row + 1 < H) b |= m1[w + Wp];
            if (row + 2 < H) b |= m1[w + 2 * Wp];
            if (row + 3 < H) b |= m1[w + 3 * Wp];
            m4[w] = b;
        }
    __syncthreads();
    if (depth)
        for (int w = tid; w < H * Wp; w += NT) {
            int row = w / Wp;
            uint32_t b = m4[w];
            if (row + 4 < H) b |= m4[w + 4 * Wp];
            if (row + 8 < H) b |= m4[w + 8 * Wp];
            if (row + 12 < H) b |= m4[w + 12 * Wp];
            m16[w] = b;
        }

    // ---- 2. lattice planes are projected once each and kept two at a time
    const float* P = pose + (size_t)view * 16;
    const float p00 = P[0], p01 = P[1], p02 = P[2], p03 = P[3];
    const float p10 = P[4], p11 = P[5], p12 = P[6], p13 = P[7];
    const float p20 = P[8], p21 = P[9], p22 = P[10], p23 = P[11];
    // cam = p0*wx + p1*wy + p2*wz + p3 is accumulated in that order (one FMA per term), so the first two terms do
    // not depend on the plane: they are formed once per view with the same operations
    {
        const float rn = 1.0f / (float)n;
        for (int q = tid; q < nn; q += NT) {
            const int ly = (int)(((float)q + 0.5f) * rn), lx = q - ly * n;      // exact quotient for q < 2^20
            // grid_to_world @ corner: inv*x + (-0.5)*1 (raytracing.py:35-40,64)
            const float wx = __fadd_rn(__fmul_rn(inv_dim, (float)lx), -0.5f);
            const float wy = __fadd_rn(__fmul_rn(inv_dim, (float)ly), -0.5f);
            axy[q] = fmaf(p01, wy, __fmul_rn(p00, wx));
            axy[nn + q] = fmaf(p11, wy, __fmul_rn(p10, wx));
            axy[2 * nn + q] = fmaf(p21, wy, __fmul_rn(p20, wx));
        }
    }
    auto project_plane = [&](int lz, uint32_t* out) {
        const float wz = __fadd_rn(__fmul_rn(inv_dim, (float)lz), -0.5f);
        for (int q = tid; q < nn; q += NT) {
            const float camx = fmaf(p03, 1.f, fmaf(p02, wz, axy[q]));
            const float camy = fmaf(p13, 1.f, fmaf(p12, wz, axy[nn + q]));
            const float camz = fmaf(p23, 1.f, fmaf(p22, wz, axy[2 * nn + q]));
            const float az = fabsf(camz);
            const float inv_az = rcp_approx(az);
            const int u = project_px(__fmul_rn(camx, focal), az, inv_az, cx, clamp_max);     // raytracing.py:71
            const int v = project_px(__fmul_rn(-camy, focal), az, inv_az, cy, clamp_max);    // :70,72
            out[q] = (uint32_t)u | ((uint32_t)v << 16);
        }
    };
    // per cell (x, y) of a plane: packed (umin | vmin << 16, umax | vmax << 16) over its 4 corners. A cell is
    // always handled by the same thread (slice after slice), so the previous plane's value needs no barrier.
    const float rdim = 1.0f / (float)dim;
    auto cell_quad = [&](const uint32_t* in, int q) {
        const int y = (int)(((float)q + 0.5f) * rdim), x = q - y * dim;          // exact quotient for q < 2^20
        const uint32_t* c = in + y * n + x;
        const uint32_t a = c[0], b = c[1], e = c[n], f = c[n + 1];
        return make_uint2(__vminu2(__vminu2(a, b), __vminu2(e, f)), __vmaxu2(__vmaxu2(a, b), __vmaxu2(e, f)));
    };
    __syncthreads();
    project_plane(z0, lat);
    __syncthreads();
    for (int q = tid; q < d2; q += NT) quad[q] = cell_quad(lat, q);

    // ---- 3. slice by slice: project plane z+1, then test the voxels between planes z and z+1
    const int nvox = d2 * dim;                      // dim <= 1023: fits
    const bool aligned = (d2 & 31) == 0;
    uint32_t* const ow = occ_bits ? occ_bits + (size_t)view * ((nvox + 31) >> 5) : nullptr;
    float* const of = occ_f32 ? occ_f32 + (size_t)view * nvox : nullptr;
    long long* const im = index_map ? index_map + (size_t)view * 4 * nvox : nullptr;   // raytracing.py:76-80
    for (int z = z0; z < z1; ++z) {
        const int cur = (z - z0) & 1, nxt = cur ^ 1;
        project_plane(z + 1, lat + nxt * nn);
        __syncthreads();                            // plane z+1 projected; plane z-1's buffer is free again
        const uint32_t* lnx = lat + nxt * nn;
        for (int base = tid - lane; base < d2; base += NT) {
            const int q = base + lane;
            const int lin = z * d2 + q;
            bool occ = false;
            if (q < d2) {
                const uint2 A = quad[q], B = cell_quad(lnx, q);
                quad[q] = B;
                const uint32_t mn = __vminu2(A.x, B.x), mx = __vmaxu2(A.y, B.y);
                if (im) {
                    im[lin] = mn & 0xffffu; im[nvox + lin] = mn >> 16;
                    im[2 * (size_t)nvox + lin] = mx & 0xffffu; im[3 * (size_t)nvox + lin] = mx >> 16;
                }
                // depth[vmin:vmax, umin:umax] > 0 with numpy's end clipping (raytracing.py:134): clip both packed
                // corners to (W, H) at once
                const uint32_t lim = (uint32_t)W | ((uint32_t)H << 16);
                const uint32_t lo = __vminu2(mn, lim), hi = __vminu2(mx, lim);
                const int u0 = lo & 0xffffu, v0 = lo >> 16, u1 = hi & 0xffffu, v1 = hi >> 16;
                if (depth && u1 > u0 && v1 > v0) {
                    const int wa = u0 >> 5, wb = (u1 - 1) >> 5;
                    const uint32_t ma = 0xffffffffu << (u0 & 31);
                    const uint32_t mb = 0xffffffffu >> (31 - ((u1 - 1) & 31));
                    // rows [v0, v1) are covered by at most four reads of the coarsest fitting mask: windows of
                    // h rows starting at v0, v0+h, ... and one ending at v1 (overlap is harmless for an OR)
                    const int nrow = v1 - v0;
                    const uint32_t* mk = nrow >= 16 ? m16 : (nrow >= 4 ? m4 : m1);
                    const int h = nrow >= 16 ? 16 : (nrow >= 4 ? 4 : 1);
                    const uint32_t* last = mk + (v1 - h) * Wp;
                    uint32_t hit = 0;
                    if (wb - wa <= 1 && nrow <= 4 * h) {
                        // the usual case (a voxel covers a few pixels): at most four windows and two words,
                        // read without loops; windows past the last one are clamped onto it
                        const uint32_t* r0 = mk + v0 * Wp;
                        const uint32_t* r1 = mk + min(v0 + h, v1 - h) * Wp;
                        const uint32_t* r2 = mk + min(v0 + 2 * h, v1 - h) * Wp;
                        const uint32_t a = (r0[wa] | r1[wa] | r2[wa] | last[wa]) & ma;
                        const uint32_t b = (r0[wb] | r1[wb] | r2[wb] | last[wb]) & mb;
                        hit = wa == wb ? (a & b) : (a | b);     // one word: both edge masks apply to it
                    } else {
                        for (int w = wa; w <= wb; ++w) {
                            uint32_t acc = last[w];
                            for (const uint32_t* r = mk + v0 * Wp; r < last; r += h * Wp) acc |= r[w];
                            if (w == wa) acc &= ma;
                            if (w == wb) acc &= mb;
                            hit |= acc;
                        }
                    }
                    occ = hit != 0;
                }
                if (of) of[lin] = occ ? 1.f : 0.f;
            }
            const uint32_t word = __ballot_sync(0xffffffffu, occ);
            if (ow) {
                if (aligned) { if (lane == 0) ow[lin >> 5] = word; }
                else if (occ) atomicOr(ow + (lin >> 5), 1u << (lin & 31));
            }
        }
    }
}

}  // namespace

int raycast_launch(const uint8_t* depth, const float* pose, int32_t n_view, int32_t H, int32_t W,
                   float focal, float cx, float cy, int32_t clamp_max, int32_t dim,
                   uint32_t* occ_bits, float* occ_f32, int64_t* index_map, cudaStream_t stream)
{
    if (n_view < 0 || H < 1 || W < 1 || dim < 1 || dim > 1023 || clamp_max < 0 || clamp_max > 65535) {
        set_error("raycast: bad arguments (n_view=%d H=%d W=%d dim=%d clamp_max=%d)", n_view, H, W, dim, clamp_max);
        return G3D_ERR_INVALID;
    }
    if (n_view == 0) return G3D_OK;
    if (!pose || (!depth && (occ_bits || occ_f32))) { set_error("raycast: null input"); return G3D_ERR_INVALID; }
    constexpr int NT = 512;
    const int n = dim + 1;
    const size_t smem = (size_t)H * ((W + 31) / 32) * 4 * 3 + ((size_t)2 * n * n + 1) * 4 + (size_t)dim * dim * 8 +
                        ((size_t)3 * n * n + 1) * 4 + 16;
    if (smem > 227 * 1024) {
        set_error("raycast: image %dx%d needs %zu bytes of shared memory (max 232448)", H, W, smem);
        return G3D_ERR_INVALID;
    }
    // z-range per CTA: the whole grid when the batch alone fills the GPU (the mask is then built once per
    // view), thinner slabs for small batches so that ~2 CTAs per SM exist
    int zs = dim;
    const int target = 2 * sm_count();
    while (zs > 1 && (int64_t)n_view * ((dim + zs - 1) / zs) < target) zs = (zs + 1) / 2;
    auto kern = dim == 32 ? raycast_kernel<NT, 32> : raycast_kernel<NT, 0>;
    G3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nvox = (int64_t)dim * dim * dim;
    if (occ_bits && ((dim * dim) & 31))
        G3D_CUDA_TRY(cudaMemsetAsync(occ_bits, 0, (size_t)n_view * ((nvox + 31) / 32) * 4, stream));
    dim3 grid((unsigned)n_view, (unsigned)((dim + zs - 1) / zs));
    prof_begin(stream);
    kern<<<grid, NT, smem, stream>>>(depth, pose, H, W, focal, cx, cy, clamp_max, dim, zs,
                                                   1.0f / (float)dim, occ_bits, occ_f32,
                                                   reinterpret_cast<long long*>(index_map));
    prof_mark(ST_RAYCAST, stream);
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

}  // namespace g3d

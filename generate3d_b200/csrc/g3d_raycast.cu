// g3d_raycast.cu -- depth view + pose -> 32^3 input occupancy ("visual cone") on sm_100a.
//
// Reproduces compute_index_mapping + raycast_depth (src/scripts/raytracing.py:
// 43-82, 122-144) bit for bit. The reference projects the 8 corners of every
// voxel (262,144 projections per view), takes the clamped pixel bounding box and
// then runs a 32,768-iteration Python loop with a device->host sync per voxel.
// Here one CTA handles (view, z-slab):
//
//  1. the u8 depth raster is read ONCE with 16-byte loads and reduced to a
//     `depth > 0` bit mask in shared memory (H x W/32 words; 8 KB at 256^2);
//  2. the corners are shared lattice points, so each of the (dim+1)^3 lattice
//     points is projected once (7.3x fewer projections) with the reference's
//     arithmetic: cam = pose @ (g2w @ p) accumulated k = 0..3 with one FMA per
//     term (the association MKL sgemm uses for raytracing.py:64, pinned by the
//     golden fixtures), y flipped, u = x*f/|z| + cx, round-half-even, clamp;
//  3. each voxel takes min/max over its 8 lattice points and tests its box
//     against the mask a row at a time (half-open, clipped to the image as
//     numpy slicing does); a warp ballots 32 voxels into one output word, so
//     the bit-packed grid is written without atomics.
//
// Roofline: HBM. Algorithmic bytes per view = H*W (depth, read once) + 64
// (pose) + dim^3/8 (packed occupancy) = 69,696 B at 256^2, 266,304 B at 512^2.
#include "g3d_internal.h"

namespace g3d {

namespace {

__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t w)
{
    // one bit per non-zero byte of w, in byte order
    uint32_t m = __vcmpne4(w, 0u) & 0x08040201u;
    return (m | (m >> 8) | (m >> 16) | (m >> 24)) & 0xfu;
}

// torch.round(p).long() then clamp(0, cmax) (raytracing.py:73-75). A NaN/inf/huge
// float converts to INT64_MIN on x86-64, which the clamp turns into 0.
__device__ __forceinline__ int round_clamp(float p, int cmax)
{
    float r = rintf(p);
    if (!(fabsf(r) < 9.2e18f)) return 0;
    if (r <= 0.f) return 0;
    if (r >= (float)cmax) return cmax;
    return (int)r;
}

template <int NT>
__global__ void __launch_bounds__(NT)
raycast_kernel(const uint8_t* __restrict__ depth, const float* __restrict__ pose, int H, int W,
               float focal, float cx, float cy, int clamp_max, int dim, int zs, float inv_dim,
               uint32_t* __restrict__ occ_bits, float* __restrict__ occ_f32,
               long long* __restrict__ index_map)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int view = blockIdx.x;
    const int z0 = blockIdx.y * zs;
    const int z1 = min(dim, z0 + zs);
    const int Wp = (W + 31) >> 5;
    const int n = dim + 1;
    uint32_t* mask = smem;                          // [H * Wp]
    uint32_t* uv = smem + (size_t)H * Wp;           // [(z1-z0+1) * n * n]  u | v << 16
    const uint8_t* D = depth + (size_t)view * H * W;
    const int tid = threadIdx.x;

    // ---- 1. depth > 0 bit mask (depth == NULL: index map only)
    if (!depth) {
    } else if ((W & 31) == 0 && ((reinterpret_cast<uintptr_t>(D) & 15) == 0)) {
        const uint4* D4 = reinterpret_cast<const uint4*>(D);
        uint16_t* m16 = reinterpret_cast<uint16_t*>(mask);
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
            mask[w] = b;
        }
    }

    // ---- 2. lattice projection
    const float* P = pose + (size_t)view * 16;
    const float p00 = P[0], p01 = P[1], p02 = P[2], p03 = P[3];
    const float p10 = P[4], p11 = P[5], p12 = P[6], p13 = P[7];
    const float p20 = P[8], p21 = P[9], p22 = P[10], p23 = P[11];
    const int nlat = (z1 - z0 + 1) * n * n;
    for (int q = tid; q < nlat; q += NT) {
        int lx = q % n, r = q / n;
        int ly = r % n, lz = z0 + r / n;
        // grid_to_world @ corner: inv*x + (-0.5)*1 (raytracing.py:35-40,64)
        float wx = __fadd_rn(__fmul_rn(inv_dim, (float)lx), -0.5f);
        float wy = __fadd_rn(__fmul_rn(inv_dim, (float)ly), -0.5f);
        float wz = __fadd_rn(__fmul_rn(inv_dim, (float)lz), -0.5f);
        float camx = fmaf(p03, 1.f, fmaf(p02, wz, fmaf(p01, wy, __fmul_rn(p00, wx))));
        float camy = fmaf(p13, 1.f, fmaf(p12, wz, fmaf(p11, wy, __fmul_rn(p10, wx))));
        float camz = fmaf(p23, 1.f, fmaf(p22, wz, fmaf(p21, wy, __fmul_rn(p20, wx))));
        float az = fabsf(camz);
        float u = __fadd_rn(__fdiv_rn(__fmul_rn(camx, focal), az), cx);     // raytracing.py:71
        float v = __fadd_rn(__fdiv_rn(__fmul_rn(-camy, focal), az), cy);    // :70,72
        uv[q] = (uint32_t)round_clamp(u, clamp_max) | ((uint32_t)round_clamp(v, clamp_max) << 16);
    }
    __syncthreads();

    // ---- 3. per-voxel box test
    const int d2 = dim * dim;
    const int nv = (z1 - z0) * d2;
    const long long nvox = (long long)d2 * dim;
    const long long pw = (nvox + 31) >> 5;
    const bool aligned = (d2 & 31) == 0;
    const int lane = tid & 31;
    for (int base = tid - lane; base < nv; base += NT) {
        int q = base + lane;
        bool valid = q < nv;
        bool occ = false;
        long long lin = (long long)z0 * d2 + q;
        if (valid) {
            int x = q % dim, r = q / dim;
            int y = r % dim, zl = r / dim;
            const uint32_t* c = uv + (zl * n + y) * n + x;
            uint32_t umin = 0xffffu, vmin = 0xffffu, umax = 0, vmax = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t p = c[(k & 1) + ((k >> 1) & 1) * n + (k >> 2) * n * n];
                uint32_t pu = p & 0xffffu, pv = p >> 16;
                umin = min(umin, pu); umax = max(umax, pu);
                vmin = min(vmin, pv); vmax = max(vmax, pv);
            }
            if (index_map) {
                long long* im = index_map + (long long)view * 4 * nvox + lin;   // raytracing.py:76-80
                im[0] = umin; im[nvox] = vmin; im[2 * nvox] = umax; im[3 * nvox] = vmax;
            }
            // depth[vmin:vmax, umin:umax] > 0 with numpy's end clipping (raytracing.py:134)
            int u0 = min((int)umin, W), u1 = min((int)umax, W);
            int v0 = min((int)vmin, H), v1 = min((int)vmax, H);
            if (depth && u1 > u0 && v1 > v0) {
                int wa = u0 >> 5, wb = (u1 - 1) >> 5;
                uint32_t ma = 0xffffffffu << (u0 & 31);
                uint32_t mb = 0xffffffffu >> (31 - ((u1 - 1) & 31));
                for (int row = v0; row < v1 && !occ; ++row) {
                    const uint32_t* mr = mask + row * Wp;
                    for (int w = wa; w <= wb; ++w) {
                        uint32_t m = mr[w];
                        if (w == wa) m &= ma;
                        if (w == wb) m &= mb;
                        if (m) { occ = true; break; }
                    }
                }
            }
            if (occ_f32) occ_f32[(long long)view * nvox + lin] = occ ? 1.f : 0.f;
        }
        uint32_t word = __ballot_sync(0xffffffffu, occ);
        if (occ_bits) {
            uint32_t* ow = occ_bits + (long long)view * pw;
            if (aligned) { if (lane == 0 && base < nv) ow[lin >> 5] = word; }
            else if (occ) atomicOr(ow + (lin >> 5), 1u << (lin & 31));
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
    const size_t mask_bytes = (size_t)H * ((W + 31) / 32) * 4;
    auto smem_for = [&](int zs) { return mask_bytes + (size_t)(zs + 1) * n * n * 4; };
    // z-slab depth: small enough that the grid covers the GPU ~2x and the CTA fits
    // several times per SM, large enough to amortise the per-CTA mask rebuild
    int zs = dim;
    const int target = 2 * sm_count();
    while (zs > 1 && ((int64_t)n_view * ((dim + zs - 1) / zs) < target || smem_for(zs) > 72 * 1024)) zs = (zs + 1) / 2;
    size_t smem = smem_for(zs);
    if (smem > 227 * 1024) {
        set_error("raycast: image %dx%d needs %zu bytes of shared memory (max 232448)", H, W, smem);
        return G3D_ERR_INVALID;
    }
    G3D_CUDA_TRY(cudaFuncSetAttribute(raycast_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nvox = (int64_t)dim * dim * dim;
    if (occ_bits && ((dim * dim) & 31))
        G3D_CUDA_TRY(cudaMemsetAsync(occ_bits, 0, (size_t)n_view * ((nvox + 31) / 32) * 4, stream));
    dim3 grid((unsigned)n_view, (unsigned)((dim + zs - 1) / zs));
    prof_begin(stream);
    raycast_kernel<NT><<<grid, NT, smem, stream>>>(depth, pose, H, W, focal, cx, cy, clamp_max, dim, zs,
                                                   1.0f / (float)dim, occ_bits, occ_f32,
                                                   reinterpret_cast<long long*>(index_map));
    prof_mark(ST_RAYCAST, stream);
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

}  // namespace g3d

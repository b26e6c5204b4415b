// g3d_next.cu -- the first two "next" rows of the scope table:
//   occ_to_df      occupancy -> 26-neighbour chamfer distance field (Network/data_utils.py:115-168,
//                  occ_to_df.py:16-62), used for the occ_df_gt tensor and inside validation;
//   raycast colour per-voxel mean colour of the valid-depth pixels of the voxel's box (raytracing.py:85-120).
#include "g3d_internal.h"

namespace g3d {

namespace {

// ------------------------------------------------------------------------------- occ_to_df ----
// What the reference computes, once its aliasing is taken into account (`df = new_df` makes both names
// one tensor, so the allclose loop ends after its first trip): a (dim+2)^3 grid, 0 at occupied cells and
// inf elsewhere, relaxed with the 27-offset Euclidean kernel exactly TWICE -- or not at all when the first
// relaxation moves no cell from inf to finite (empty or full grids), because then the loop is never entered
// and the untouched initial field is returned. Values above trunc become `over`.
// After one relaxation a cell can only hold 0, 1, sqrt2, sqrt3 or inf, so the intermediate grid is kept as
// one byte per cell in shared memory.
__global__ void __launch_bounds__(512)
occ_to_df_kernel(const float* __restrict__ occ, int dim, int pred, float trunc, float over, float* __restrict__ out)
{
    extern __shared__ uint8_t s_grid[];             // [2][(dim+2)^3] codes: 0 -> 0, 1 -> 1, 2 -> sqrt2, 3 -> sqrt3, 4 -> inf
    const int n = dim + 2, n2 = n * n, n3 = n2 * n, d3 = dim * dim * dim;
    uint8_t* c0 = s_grid;
    uint8_t* c1 = s_grid + n3;
    const float* in = occ + (size_t)blockIdx.x * d3;
    float* o = out + (size_t)blockIdx.x * d3;
    const float sq2 = sqrtf(2.f), sq3 = sqrtf(3.f), inf = __int_as_float(0x7f800000);
    for (int q = threadIdx.x; q < n3; q += blockDim.x) { c0[q] = 4; c1[q] = 4; }
    __syncthreads();
    for (int q = threadIdx.x; q < d3; q += blockDim.x) {
        const int x = q % dim, y = (q / dim) % dim, z = q / (dim * dim);
        const float v = in[q];
        bool on;
        if (pred) on = (1.f / (1.f + expf(-v))) >= 0.5f;            // preprocess_occ: sigmoid >= 0.5
        else on = (v == 1.f);
        if (on) c0[(z + 1) * n2 + (y + 1) * n + (x + 1)] = 0;
    }
    __syncthreads();
    int moved = 0;
    for (int q = threadIdx.x; q < d3; q += blockDim.x) {
        const int x = q % dim, y = (q / dim) % dim, z = q / (dim * dim);
        const int c = (z + 1) * n2 + (y + 1) * n + (x + 1);
        int best = 4;
        if (c0[c] == 0) best = 0;
        else {
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int k = (dz != 0) + (dy != 0) + (dx != 0);   // 1, 2, 3 -> 1, sqrt2, sqrt3
                        if (k && c0[c + dz * n2 + dy * n + dx] == 0) best = min(best, k);
                    }
            if (best < 4) moved = 1;
        }
        c1[c] = (uint8_t)best;
    }
    const int any_moved = __syncthreads_or(moved);
    const float val[5] = { 0.f, 1.f, sq2, sq3, inf };
    const float ker[4] = { 0.f, 1.f, sq2, sq3 };
    for (int q = threadIdx.x; q < d3; q += blockDim.x) {
        const int x = q % dim, y = (q / dim) % dim, z = q / (dim * dim);
        const int c = (z + 1) * n2 + (y + 1) * n + (x + 1);
        float best;
        if (!any_moved) best = c0[c] == 0 ? 0.f : inf;
        else {
            best = inf;
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int k = (dz != 0) + (dy != 0) + (dx != 0);
                        best = fminf(best, __fadd_rn(val[c1[c + dz * n2 + dy * n + dx]], ker[k]));
                    }
        }
        o[q] = best > trunc ? over : best;
    }
}

// ---------------------------------------------------------------------------- colour raycast ----
__device__ __forceinline__ int round_clamp_px(float p, int cmax)
{
    float r = rintf(p);
    if (!(fabsf(r) < 9.2e18f)) return 0;
    if (r <= 0.f) return 0;
    if (r >= (float)cmax) return cmax;
    return (int)r;
}

__global__ void __launch_bounds__(256)
raycast_color_kernel(const uint8_t* __restrict__ depth, const uint8_t* __restrict__ color, const float* __restrict__ pose,
                     int H, int W, float focal, float cx, float cy, int clamp_max, int dim, float inv_dim,
                     uint16_t* __restrict__ out)
{
    const int view = blockIdx.y;
    const int d2 = dim * dim, d3 = d2 * dim;
    const int lin = blockIdx.x * blockDim.x + threadIdx.x;              // z*dim^2 + y*dim + x (raytracing.py:44-51)
    if (lin >= d3) return;
    const int x = lin % dim, y = (lin / dim) % dim, z = lin / d2;
    const float* P = pose + (size_t)view * 16;
    int umin = 1 << 30, vmin = 1 << 30, umax = -1, vmax = -1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float wx = __fadd_rn(__fmul_rn(inv_dim, (float)(x + (k & 1))), -0.5f);
        const float wy = __fadd_rn(__fmul_rn(inv_dim, (float)(y + ((k >> 1) & 1))), -0.5f);
        const float wz = __fadd_rn(__fmul_rn(inv_dim, (float)(z + (k >> 2))), -0.5f);
        const float camx = fmaf(P[3], 1.f, fmaf(P[2], wz, fmaf(P[1], wy, __fmul_rn(P[0], wx))));
        const float camy = fmaf(P[7], 1.f, fmaf(P[6], wz, fmaf(P[5], wy, __fmul_rn(P[4], wx))));
        const float camz = fmaf(P[11], 1.f, fmaf(P[10], wz, fmaf(P[9], wy, __fmul_rn(P[8], wx))));
        const float az = fabsf(camz);
        const int u = round_clamp_px(__fadd_rn(__fdiv_rn(__fmul_rn(camx, focal), az), cx), clamp_max);
        const int v = round_clamp_px(__fadd_rn(__fdiv_rn(__fmul_rn(-camy, focal), az), cy), clamp_max);
        umin = min(umin, u); umax = max(umax, u); vmin = min(vmin, v); vmax = max(vmax, v);
    }
    const int u0 = min(umin, W), u1 = min(umax, W), v0 = min(vmin, H), v1 = min(vmax, H);
    const uint8_t* D = depth + (size_t)view * H * W;
    const uint8_t* C = color + (size_t)view * H * W * 3;
    unsigned sr = 0, sg = 0, sb = 0, cnt = 0;
    for (int v = v0; v < v1; ++v)
        for (int u = u0; u < u1; ++u)
            if (D[(size_t)v * W + u] > 0) {                             // raytracing.py:104-105
                const uint8_t* c = C + ((size_t)v * W + u) * 3;
                sr += c[0]; sg += c[1]; sb += c[2]; ++cnt;
            }
    // torch.mean of <= a few hundred byte values: the sum is exact in float, the mean is sum / count;
    // then ceil and uint16 (raytracing.py:107-111); 256 marks "no valid pixel" (:96)
    uint16_t r = 256, g = 256, b = 256;
    if (cnt) {
        r = (uint16_t)ceilf(__fdiv_rn((float)sr, (float)cnt));
        g = (uint16_t)ceilf(__fdiv_rn((float)sg, (float)cnt));
        b = (uint16_t)ceilf(__fdiv_rn((float)sb, (float)cnt));
    }
    uint16_t* o = out + ((size_t)view * d3 + ((size_t)x * dim + y) * dim + z) * 3;   // [x][y][z][rgb] (:111)
    o[0] = r; o[1] = g; o[2] = b;
}

}  // namespace

}  // namespace g3d

using namespace g3d;

extern "C" int g3d_occ_to_df_batch(const float* occ, int32_t n, int32_t dim, int32_t pred, float trunc, float over,
                                   float* out, void* cuda_stream)
{
    if (n < 0 || dim < 1 || dim > 56) { set_error("occ_to_df: bad n/dim (dim <= 56)"); return G3D_ERR_INVALID; }
    if (n == 0) return G3D_OK;
    if (!occ || !out) { set_error("occ_to_df: NULL pointer"); return G3D_ERR_INVALID; }
    const size_t smem = (size_t)2 * (dim + 2) * (dim + 2) * (dim + 2);
    G3D_CUDA_TRY(cudaFuncSetAttribute(occ_to_df_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    occ_to_df_kernel<<<(unsigned)n, 512, smem, (cudaStream_t)cuda_stream>>>(occ, dim, pred, trunc, over, out);
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

extern "C" int g3d_raycast_color_batch(const uint8_t* depth, const uint8_t* color, const float* pose, int32_t n_view,
                                       int32_t H, int32_t W, float focal, float cx, float cy, int32_t clamp_max,
                                       int32_t dim, uint16_t* out, void* cuda_stream)
{
    if (n_view < 0 || H < 1 || W < 1 || dim < 1 || dim > 1023 || clamp_max < 0 || n_view > 65535) {
        set_error("raycast_color: bad arguments");
        return G3D_ERR_INVALID;
    }
    if (n_view == 0) return G3D_OK;
    if (!depth || !color || !pose || !out) { set_error("raycast_color: NULL pointer"); return G3D_ERR_INVALID; }
    const int d3 = dim * dim * dim;
    dim3 grid((unsigned)((d3 + 255) / 256), (unsigned)n_view);
    raycast_color_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(depth, color, pose, H, W, focal, cx, cy, clamp_max,
                                                                      dim, 1.0f / (float)dim, out);
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

// ------------------------------------------------------------------------ per-pixel ray march ----
// the older raytracing_approach1.py:13-51. Every pixel shoots a ray from the
// z_near plane through a camera-aligned voxel box and marks the voxels it crosses in half-voxel steps; a voxel
// keeps the colour of the LAST pixel (u outer, v inner loop) that touched it. The reference does this in float64
// numpy scalars, so the march is replayed in double precision with explicitly rounded operations; "last writer"
// becomes an atomicMax over u*H+v per voxel.
namespace g3d {
namespace {

__global__ void __launch_bounds__(256)
raymarch_kernel(const uint8_t* __restrict__ depth, int H, int W, double cx, double cy, double focal, double z_near,
                double grid_size, int dim, uint32_t* __restrict__ last_px)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= H * W) return;
    const int u = p / H, v = p - u * H;              // u outer, v inner (raytracing_approach1.py:29-30)
    if (depth[(size_t)v * W + u] == 0) return;       // :45 -- such a pixel never writes
    const double vs = grid_size / (double)dim;       // voxel_grid.py:14
    const double minx = -grid_size / 2, miny = -grid_size / 2, minz = -z_near;           // voxel_grid.py:167-168
    const double maxx = __dadd_rn(minx, grid_size), maxy = __dadd_rn(miny, grid_size);
    const double maxz = __dsub_rn(__dadd_rn(minz, grid_size), __dmul_rn(2.0, grid_size)); // voxel_grid.py:16-19
    const double x = __ddiv_rn(__dmul_rn((double)u - cx, z_near), focal);                // :31-32
    const double y = __ddiv_rn(__dmul_rn((double)v - cy, z_near), focal);
    double rx = x, ry = -y, rz = -z_near;
    double len = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz)));   // :34
    const double ux = __ddiv_rn(rx, len), uy = __ddiv_rn(ry, len), uz = __ddiv_rn(rz, len);
    const double half = vs / 2;
    for (int guard = 0; guard < (1 << 20); ++guard) {
        if (!(minx <= rx && rx <= maxx && miny <= ry && ry <= maxy && maxz <= rz && rz <= minz)) break;   // voxel_grid.py:55-63
        const int gi = (int)__ddiv_rn(__dsub_rn(rx, minx), vs);                          // voxel_grid.py:42-47
        const int gj = (int)__ddiv_rn(__dsub_rn(ry, miny), vs);
        const int gk = (int)__ddiv_rn(-__dsub_rn(rz, minz), vs);
        if (gi <= dim - 1 && gj <= dim - 1 && gk <= dim - 1)                             // voxel_grid.py:49-50
            atomicMax(last_px + ((size_t)gi * dim + gj) * dim + gk, (uint32_t)p + 1u);
        len = __dadd_rn(len, half);                                                      // :50-51
        rx = __dmul_rn(ux, len); ry = __dmul_rn(uy, len); rz = __dmul_rn(uz, len);
    }
}

__global__ void raymarch_resolve_kernel(const uint32_t* __restrict__ last_px, const uint8_t* __restrict__ color, int H, int W,
                                        int n, uint8_t* __restrict__ occ, uint8_t* __restrict__ rgb)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint32_t k = last_px[q];
    occ[q] = k ? 1 : 0;
    uint8_t r = 0, g = 0, b = 0;
    if (k) {
        const int p = (int)(k - 1), u = p / H, v = p - u * H;
        const uint8_t* c = color + ((size_t)v * W + u) * 3;
        r = c[0]; g = c[1]; b = c[2];
    }
    rgb[3 * q] = r; rgb[3 * q + 1] = g; rgb[3 * q + 2] = b;
}

}  // namespace
}  // namespace g3d

extern "C" int g3d_raymarch(const uint8_t* depth, const uint8_t* color, int32_t H, int32_t W, double cx, double cy,
                            double focal, double z_near, double z_far, int32_t dim, uint8_t* occ, uint8_t* rgb,
                            uint32_t* scratch, void* cuda_stream)
{
    if (H < 1 || W < 1 || dim < 1 || dim > 1024 || !(focal > 0) || !(z_near > 0)) { set_error("raymarch: bad arguments"); return G3D_ERR_INVALID; }
    if (!depth || !color || !occ || !rgb || !scratch) { set_error("raymarch: NULL pointer"); return G3D_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const int n = dim * dim * dim;
    const double grid_size = fabs(z_far - z_near);                                       // voxel_grid.py:166
    G3D_CUDA_TRY(cudaMemsetAsync(scratch, 0, (size_t)n * 4, s));
    raymarch_kernel<<<(H * W + 255) / 256, 256, 0, s>>>(depth, H, W, cx, cy, focal, z_near, grid_size, dim, scratch);
    raymarch_resolve_kernel<<<(n + 255) / 256, 256, 0, s>>>(scratch, color, H, W, n, occ, rgb);
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

// ------------------------------------------------------------------ z-buffered index map ----
// Network/perspective_projection.py:65-128 (ProjectionHelper.compute_index_mapping): every occupied voxel is
// projected (8 corners, pixel bounding box half-open as numpy slices it, depth = min |z_cam| of the corners) and
// painted into the image in increasing voxel order, a pixel changing hands only for a strictly smaller depth.
// The result per pixel is therefore the minimum of (depth, voxel index) over the voxels whose box covers it --
// an atomicMin on depth_bits << 32 | index, in any order. The int64 output doubles as the key buffer.
#include "g3d_project.cuh"

namespace g3d {
namespace {

__global__ void __launch_bounds__(256)
zbuffer_paint_kernel(const uint32_t* __restrict__ occ_bits, const float* __restrict__ pose, int n_view, int H, int W,
                     float focal, float cx, float cy, int clamp_max, int dim, float inv_dim,
                     unsigned long long* __restrict__ key)
{
    const int gv = blockIdx.y;                                  // grid * n_view + view
    const int grid = gv / n_view;
    const int nvox = dim * dim * dim;
    const uint32_t* bits = occ_bits + (size_t)grid * ((nvox + 31) >> 5);
    const float* P = pose + (size_t)gv * 16;
    unsigned long long* K = key + (size_t)gv * H * W;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int w0 = warp; w0 < (nvox + 31) >> 5; w0 += nwarp) {
        uint32_t m = __ldg(bits + w0);
        while (m) {
            const int lin = (w0 << 5) + __ffs(m) - 1;           // lin = x + dim*(y + dim*z) (perspective_projection.py:78-84)
            m &= m - 1;
            if (lin >= nvox) break;
            const int z = lin / (dim * dim), r = lin - z * dim * dim, y = r / dim, x = r - y * dim;
            // lanes 0..7 project one corner each (vertex_coords, :88-91: bit 0 -> x, bit 1 -> y, bit 2 -> z)
            const int c = lane & 7;
            const float wx = __fadd_rn(__fmul_rn(inv_dim, (float)(x + (c & 1))), -0.5f);
            const float wy = __fadd_rn(__fmul_rn(inv_dim, (float)(y + ((c >> 1) & 1))), -0.5f);
            const float wz = __fadd_rn(__fmul_rn(inv_dim, (float)(z + (c >> 2))), -0.5f);
            // pose @ (grid_to_world @ corner), accumulated k = 0..3 with one FMA per term (as in g3d_raycast.cu)
            const float camx = fmaf(P[3], 1.f, fmaf(P[2], wz, fmaf(P[1], wy, __fmul_rn(P[0], wx))));
            const float camy = fmaf(P[7], 1.f, fmaf(P[6], wz, fmaf(P[5], wy, __fmul_rn(P[4], wx))));
            const float camz = fmaf(P[11], 1.f, fmaf(P[10], wz, fmaf(P[9], wy, __fmul_rn(P[8], wx))));
            const float az = fabsf(camz);
            const float inv_az = rcp_approx(az);
            int u = project_px(__fmul_rn(camx, focal), az, inv_az, cx, clamp_max);          // :104-110
            int v = project_px(__fmul_rn(-camy, focal), az, inv_az, cy, clamp_max);
            int umin = u, umax = u, vmin = v, vmax = v;
            float depth = az;                                    // torch.min over the 8 corners (:98) propagates NaN
            bool bad = !(az == az);
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                umin = min(umin, __shfl_xor_sync(0xffffffffu, umin, d)); umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, d));
                vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, d)); vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
                const float o = __shfl_xor_sync(0xffffffffu, depth, d);
                depth = o < depth ? o : depth;
                bad |= __shfl_xor_sync(0xffffffffu, (int)bad, d) != 0;
            }
            if (bad) continue;                                   // a NaN depth compares false everywhere: the voxel is never painted
            const int u0 = min(umin, W), u1 = min(umax, W), v0 = min(vmin, H), v1 = min(vmax, H);   // slice end clipping
            const int bw = u1 - u0, bh = v1 - v0;
            if (bw <= 0 || bh <= 0) continue;
            const unsigned long long k = ((unsigned long long)__float_as_uint(depth) << 32) | (uint32_t)lin;
            for (int q = lane; q < bw * bh; q += 32) {
                const int pv = v0 + q / bw, pu = u0 + q - (q / bw) * bw;
                atomicMin(K + (size_t)pv * W + pu, k);
            }
        }
    }
}

__global__ void zbuffer_resolve_kernel(long long* __restrict__ index_map, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = (unsigned long long)index_map[i];
    index_map[i] = (k == ~0ull) ? -1ll : (long long)(uint32_t)k;  // :121-126
}

}  // namespace
}  // namespace g3d

extern "C" int g3d_zbuffer_index_map(const uint32_t* occ_bits, const float* pose, int32_t n_grid, int32_t n_view,
                                     int32_t H, int32_t W, float focal, float cx, float cy, int32_t clamp_max,
                                     int32_t dim, int64_t* index_map, void* cuda_stream)
{
    if (n_grid < 0 || n_view < 0 || H < 1 || W < 1 || dim < 1 || dim > 1023 || clamp_max < 0 || clamp_max > 65535) {
        set_error("zbuffer_index_map: bad arguments (n_grid=%d n_view=%d H=%d W=%d dim=%d clamp_max=%d)", n_grid, n_view, H, W, dim, clamp_max);
        return G3D_ERR_INVALID;
    }
    if (n_grid == 0 || n_view == 0) return G3D_OK;
    if (!occ_bits || !pose || !index_map) { set_error("zbuffer_index_map: NULL pointer"); return G3D_ERR_INVALID; }
    if ((int64_t)n_grid * n_view > 65535) { set_error("zbuffer_index_map: at most 65535 (grid, view) pairs per call"); return G3D_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const size_t n = (size_t)n_grid * n_view * H * W;
    G3D_CUDA_TRY(cudaMemsetAsync(index_map, 0xff, n * 8, s));   // all ones: "no voxel yet", and -1 once resolved
    const int nword = (dim * dim * dim + 31) / 32;
    dim3 grid((unsigned)((nword + 7) / 8 < 64 ? (nword + 7) / 8 : 64), (unsigned)(n_grid * n_view));
    zbuffer_paint_kernel<<<grid, 256, 0, s>>>(occ_bits, pose, n_view, H, W, focal, cx, cy, clamp_max, dim,
                                              1.0f / (float)dim, reinterpret_cast<unsigned long long*>(index_map));
    zbuffer_resolve_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<long long*>(index_map), n);
    G3D_CUDA_TRY(cudaGetLastError());
    return G3D_OK;
}

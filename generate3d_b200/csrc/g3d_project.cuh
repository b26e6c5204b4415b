// g3d_project.cuh -- the reference's pixel projection of a lattice point, shared by the raycast kernels and the
// z-buffered index map: torch.round(p).long().clamp(0, cmax) of u = (x*f)/|z| + c (raytracing.py:70-75,
// Network/perspective_projection.py:104-110).
#pragma once
#include <cuda_runtime.h>

namespace g3d {

// torch.round(p).long() then clamp(0, cmax). A NaN/inf/huge float converts to INT64_MIN on x86-64, which the
// clamp turns into 0.
__device__ __forceinline__ int round_clamp(float p, int cmax)
{
    float r = rintf(p);
    if (!(fabsf(r) < 9.2e18f)) return 0;
    if (r <= 0.f) return 0;
    if (r >= (float)cmax) return cmax;
    return (int)r;
}

// u = (x*f)/|z| + c rounded half-to-even and clamped, as the reference computes it, but the IEEE division is only
// paid when it can matter. The estimate num * rcp(|z|) + c (approximate reciprocal, 1 ulp, plus one FMA rounding)
// is within |num/az| * 2^-22 + ulp(u) < 7e-4 px of the exactly rounded u for |u| < 2048, and rint() of the two can
// only differ when the estimate lies that close to a half-integer. Everything else -- non-finite values, large
// magnitudes, near-ties (0.2 % of points) -- takes the exact path.
__device__ __forceinline__ int project_px(float num, float az, float inv_az, float c, int cmax)
{
    const float est = fmaf(num, inv_az, c);
    const float r = rintf(est);
    if (fabsf(est) < 2048.f && fabsf(est - r) < 0.499f)                   // false for NaN
        return (int)fminf(fmaxf(r, 0.f), (float)cmax);
    return round_clamp(__fadd_rn(__fdiv_rn(num, az), c), cmax);
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

}  // namespace g3d

// g3d_ptd.cuh -- reference-faithful point/segment/triangle distances on the device.
//
// The CPU reference (src/dfgen/makelevelset3.cpp:4-41) is compiled for baseline
// x86-64: every float operation is a separately rounded IEEE binary32 op, no
// FMA. The sweep compares these distances with strict `<` and propagates the
// winner's triangle index, so one differing ulp changes the field downstream.
// Every operation here is therefore an explicit round-to-nearest intrinsic
// (never contracted by nvcc), division and square root included.
//
// (float)(double(a)/double(b)) of makelevelset3.cpp:9 equals the correctly
// rounded binary32 quotient (53 >= 2*24+2, double rounding is innocuous), so
// __fdiv_rn reproduces it without touching the FP64 pipe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace g3d {

struct V3 { float x, y, z; };

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return { sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z) }; }
// vec.h:327-333 / :198-204: left-to-right sums
__device__ __forceinline__ float dot3(V3 a, V3 b) { return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z)); }
// vec.h:210-220
__device__ __forceinline__ float dist3(V3 a, V3 b)
{
    V3 t = sub3(a, b);
    return __fsqrt_rn(dot3(t, t));
}
// std::min / std::max keep their NaN behaviour: min(a,b) = b<a ? b : a
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }

// makelevelset3.cpp:4-17
__device__ __forceinline__ float point_segment_distance(V3 x0, V3 x1, V3 x2)
{
    V3 e = sub3(x2, x1);
    float m2 = dot3(e, e);
    float s12 = __fdiv_rn(dot3(sub3(x2, x0), e), m2);
    if (s12 < 0.f) s12 = 0.f;
    else if (s12 > 1.f) s12 = 1.f;
    float om = sub(1.f, s12);
    V3 c = { add(mul(x1.x, s12), mul(x2.x, om)), add(mul(x1.y, s12), mul(x2.y, om)),
             add(mul(x1.z, s12), mul(x2.z, om)) };
    return dist3(x0, c);
}

// makelevelset3.cpp:20-41. The three edge cases are folded into one selection
// of two segments so that a warp whose lanes sit in different Voronoi regions
// of their triangles still runs a single instruction stream.
__device__ __forceinline__ float point_triangle_distance(V3 x0, V3 x1, V3 x2, V3 x3)
{
    V3 x13 = sub3(x1, x3), x23 = sub3(x2, x3), x03 = sub3(x0, x3);
    float m13 = dot3(x13, x13), m23 = dot3(x23, x23), d = dot3(x13, x23);
    float invdet = __fdiv_rn(1.f, std_max(sub(mul(m13, m23), mul(d, d)), 1e-30f));
    float a = dot3(x13, x03), b = dot3(x23, x03);
    float w23 = mul(invdet, sub(mul(m23, a), mul(d, b)));
    float w31 = mul(invdet, sub(mul(m13, b), mul(d, a)));
    float w12 = sub(sub(1.f, w23), w31);
    if (w23 >= 0.f && w31 >= 0.f && w12 >= 0.f) {
        V3 c = { add(add(mul(x1.x, w23), mul(x2.x, w31)), mul(x3.x, w12)),
                 add(add(mul(x1.y, w23), mul(x2.y, w31)), mul(x3.y, w12)),
                 add(add(mul(x1.z, w23), mul(x2.z, w31)), mul(x3.z, w12)) };
        return dist3(x0, c);
    }
    // w23>0: edges (1,2),(1,3); else w31>0: (1,2),(2,3); else: (1,3),(2,3)
    bool p = w23 > 0.f, q = w31 > 0.f;
    V3 a2 = (p || q) ? x2 : x3;          // first segment is x1-a2
    V3 b1 = p ? x1 : x2;                 // second segment is b1-x3
    float s1 = point_segment_distance(x0, x1, a2);
    float s2 = point_segment_distance(x0, b1, x3);
    return std_min(s1, s2);
}

// (int)x the way x86-64 cvttsd2si does it: NaN / out of range -> INT_MIN.
__device__ __forceinline__ int x86_d2i(double x)
{
    if (!(x > -2147483649.0 && x < 2147483648.0)) return INT32_MIN;
    return __double2int_rz(x);
}

}  // namespace g3d

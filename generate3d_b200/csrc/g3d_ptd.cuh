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

// ---- the same two functions on a pre-digested triangle record -------------------
// Everything that depends on the triangle alone (|x13|^2, |x23|^2, x13.x23, the
// reciprocal determinant and the three squared edge lengths) is computed once per
// triangle by tdf_prep with the SAME sequence of rounded operations, so using the
// stored values is bit-identical to recomputing them per evaluation.
struct TriRec {
    V3 x1, x2, x3;
    float m13, m23, d, invdet;      // makelevelset3.cpp:24-25
    float m2_12, m2_13, m2_23;      // mag2 of the edge vectors, makelevelset3.cpp:7
};

__device__ __forceinline__ TriRec make_trirec(V3 x1, V3 x2, V3 x3)
{
    TriRec r;
    r.x1 = x1; r.x2 = x2; r.x3 = x3;
    V3 x13 = sub3(x1, x3), x23 = sub3(x2, x3);
    r.m13 = dot3(x13, x13);
    r.m23 = dot3(x23, x23);
    r.d = dot3(x13, x23);
    r.invdet = __fdiv_rn(1.f, std_max(sub(mul(r.m13, r.m23), mul(r.d, r.d)), 1e-30f));
    V3 e12 = sub3(x2, x1), e13 = sub3(x3, x1), e23 = sub3(x3, x2);
    r.m2_12 = dot3(e12, e12);
    r.m2_13 = dot3(e13, e13);
    r.m2_23 = dot3(e23, e23);
    return r;
}

// 4 x float4 in memory: (x1, m13) (x2, m23) (x3, d) (invdet, m2_12, m2_13, m2_23)
__device__ __forceinline__ void store_trirec(float4* p, const TriRec& r)
{
    p[0] = make_float4(r.x1.x, r.x1.y, r.x1.z, r.m13);
    p[1] = make_float4(r.x2.x, r.x2.y, r.x2.z, r.m23);
    p[2] = make_float4(r.x3.x, r.x3.y, r.x3.z, r.d);
    p[3] = make_float4(r.invdet, r.m2_12, r.m2_13, r.m2_23);
}

__device__ __forceinline__ TriRec load_trirec(const float4* __restrict__ p)
{
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), e = __ldg(p + 3);
    TriRec r;
    r.x1 = { a.x, a.y, a.z }; r.x2 = { b.x, b.y, b.z }; r.x3 = { c.x, c.y, c.z };
    r.m13 = a.w; r.m23 = b.w; r.d = c.w;
    r.invdet = e.x; r.m2_12 = e.y; r.m2_13 = e.z; r.m2_23 = e.w;
    return r;
}

// point_segment_distance with the squared edge length supplied, returning the SQUARED distance (the argument of
// the final sqrt of vec.h:210-220)
__device__ __forceinline__ float point_segment_dist2_m2(V3 x0, V3 x1, V3 x2, float m2)
{
    V3 e = sub3(x2, x1);
    float s12 = __fdiv_rn(dot3(sub3(x2, x0), e), m2);
    if (s12 < 0.f) s12 = 0.f;
    else if (s12 > 1.f) s12 = 1.f;
    float om = sub(1.f, s12);
    V3 c = { add(mul(x1.x, s12), mul(x2.x, om)), add(mul(x1.y, s12), mul(x2.y, om)),
             add(mul(x1.z, s12), mul(x2.z, om)) };
    V3 t = sub3(x0, c);
    return dot3(t, t);
}

// One square root per evaluation instead of up to two: the correctly rounded sqrt is non-decreasing, so
// std::min(sqrt(q1), sqrt(q2)) == sqrt(std::min(q1, q2)) bit for bit, NaN operands included (both forms
// keep the FIRST operand unless the second compares smaller).
__device__ __forceinline__ float point_triangle_distance(V3 x0, const TriRec& t)
{
    V3 x13 = sub3(t.x1, t.x3), x23 = sub3(t.x2, t.x3), x03 = sub3(x0, t.x3);
    float a = dot3(x13, x03), b = dot3(x23, x03);
    float w23 = mul(t.invdet, sub(mul(t.m23, a), mul(t.d, b)));
    float w31 = mul(t.invdet, sub(mul(t.m13, b), mul(t.d, a)));
    float w12 = sub(sub(1.f, w23), w31);
    float q2;
    if (w23 >= 0.f && w31 >= 0.f && w12 >= 0.f) {
        V3 c = { add(add(mul(t.x1.x, w23), mul(t.x2.x, w31)), mul(t.x3.x, w12)),
                 add(add(mul(t.x1.y, w23), mul(t.x2.y, w31)), mul(t.x3.y, w12)),
                 add(add(mul(t.x1.z, w23), mul(t.x2.z, w31)), mul(t.x3.z, w12)) };
        V3 e = sub3(x0, c);
        q2 = dot3(e, e);
    } else {
        bool p = w23 > 0.f, q = w31 > 0.f;
        V3 a2 = (p || q) ? t.x2 : t.x3;           // first segment x1-a2
        float ma = (p || q) ? t.m2_12 : t.m2_13;
        V3 b1 = p ? t.x1 : t.x2;                  // second segment b1-x3
        float mb = p ? t.m2_13 : t.m2_23;
        float s1 = point_segment_dist2_m2(x0, t.x1, a2, ma);
        float s2 = point_segment_dist2_m2(x0, b1, t.x3, mb);
        q2 = std_min(s1, s2);
    }
    return __fsqrt_rn(q2);
}

// (int)x the way x86-64 cvttsd2si does it: NaN / out of range -> INT_MIN.
__device__ __forceinline__ int x86_d2i(double x)
{
    if (!(x > -2147483649.0 && x < 2147483648.0)) return INT32_MIN;
    return __double2int_rz(x);
}

}  // namespace g3d

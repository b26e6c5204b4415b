// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-linkage doorway into the reference's own distance-field sources, compiled
// IN PLACE from /root/reference (never copied into this repo): this file pulls
// in src/dfgen/makelevelset3.cpp textually so the `static` helpers are
// reachable, and exports thin wrappers. Built only by oracle/Makefile into
// oracle/_ref/libg3d_ref.so (git-ignored; travels to the GPU box).
//
// The stock src/dfgen/main.cpp cannot be compiled here (it needs Eigen, which
// is not installed), so its driver arithmetic is covered by the restatement in
// g3d_oracle.c; what this shim pins is the arithmetic core.
#include G3D_REF_MAKELEVELSET3_CPP   // -D'G3D_REF_MAKELEVELSET3_CPP="/root/reference/src/dfgen/makelevelset3.cpp"'

#include <cstdint>
#include <cstring>

extern "C" {

float g3d_ref_point_triangle_distance(const float* x0, const float* x1, const float* x2,
                                      const float* x3)
{
    return point_triangle_distance(Vec3f(x0[0], x0[1], x0[2]), Vec3f(x1[0], x1[1], x1[2]),
                                   Vec3f(x2[0], x2[1], x2[2]), Vec3f(x3[0], x3[1], x3[2]));
}

float g3d_ref_point_segment_distance(const float* x0, const float* x1, const float* x2)
{
    return point_segment_distance(Vec3f(x0[0], x0[1], x0[2]), Vec3f(x1[0], x1[1], x1[2]),
                                  Vec3f(x2[0], x2[1], x2[2]));
}

// make_level_set3 (makelevelset3.h:12-14) on flat arrays. phi_out is ni*nj*nk
// floats, x fastest (array3.h:68-78).
void g3d_ref_make_level_set3(const uint32_t* tri, int64_t ntri, const float* x, int64_t nvert,
                             const float* origin, float dx, int ni, int nj, int nk,
                             float* phi_out)
{
    std::vector<Vec3ui> faces((size_t)ntri);
    for (int64_t t = 0; t < ntri; ++t) faces[t] = Vec3ui(tri[3 * t], tri[3 * t + 1], tri[3 * t + 2]);
    std::vector<Vec3f> verts((size_t)nvert);
    for (int64_t v = 0; v < nvert; ++v) verts[v] = Vec3f(x[3 * v], x[3 * v + 1], x[3 * v + 2]);
    Array3f phi;
    make_level_set3(faces, verts, Vec3f(origin[0], origin[1], origin[2]), dx, ni, nj, nk, phi);
    std::memcpy(phi_out, &phi.a[0], sizeof(float) * (size_t)ni * nj * nk);
}

}  // extern "C"

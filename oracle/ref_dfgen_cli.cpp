// oracle/ref_dfgen_cli.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The closest thing to the reference's dfgen executable that can be built in this image: src/dfgen/main.cpp itself
// needs Eigen (through LoaderMesh.h / LoaderVOX.h), which is not installed. This file compiles the reference's OWN
// distance-field core (makelevelset3.cpp) and OWN OBJ reader (tiny_obj_loader.h) in place from /root/reference and
// restates only the Eigen-carrying glue: the argument handling and grid arithmetic of main.cpp:45-120 (with the
// reference's own Vec3f type, so the float operations are the same ones), the gather loop of LoaderMesh.h:17-50 and
// save_vox (LoaderVOX.h:48-63). Built by oracle/Makefile into oracle/_ref/dfgen_ref; used (a) to pin the product CLI's
// .df bytes to a whole-program run of reference code and (b) as the per-model wall-clock baseline of bench.py
// (process start + OBJ parse + make_level_set3 + file write, what CADVoxelization.py:25-27 pays per model).
#include G3D_REF_MAKELEVELSET3_CPP
#define TINYOBJLOADER_IMPLEMENTATION
#include G3D_REF_TINYOBJ_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <sstream>
#include <string>

int main(int argc, char* argv[])
{
    if (argc != 6) { std::printf("usage: dfgen_ref <mesh.obj> <dim> <padding> <is_unit_cube> <out.df>\n"); std::exit(-1); }
    std::string filename(argv[1]), filename_out(argv[5]);
    int dim, padding, is_unit_cube;
    { std::stringstream a(argv[2]); a >> dim; }
    { std::stringstream a(argv[3]); a >> padding; }
    { std::stringstream a(argv[4]); a >> is_unit_cube; }
    float dx = 1.0 / (dim - 2 * padding);                                              // main.cpp:57

    tinyobj::attrib_t attrib;                                                          // LoaderMesh.h:17-24
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    std::string err;
    std::string materialDir = filename.substr(0, filename.rfind("/") + 1);
    bool ret = tinyobj::LoadObj(&attrib, &shapes, &materials, &err, filename.c_str(), materialDir.c_str());
    assert(ret);
    std::vector<int> Ftmp;                                                             // LoaderMesh.h:26-42
    for (size_t s = 0; s < shapes.size(); s++) {
        size_t index_offset = 0;
        for (size_t f = 0; f < shapes[s].mesh.num_face_vertices.size(); f++) {
            int fv = shapes[s].mesh.num_face_vertices[f];
            for (size_t v = 0; v < (size_t)fv; v++) Ftmp.push_back(shapes[s].mesh.indices[index_offset + v].vertex_index);
            index_offset += fv;
        }
    }
    assert(attrib.vertices.size() / 3 > 0 && Ftmp.size() / 3 > 0 && "Error loading mesh.");

    Vec3f min_box(std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()),
          max_box(-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max());
    std::vector<Vec3f> vertList;
    std::vector<Vec3ui> faceList;
    for (size_t i = 0; i < attrib.vertices.size() / 3; i++) {                          // main.cpp:71-76
        Vec3f point(attrib.vertices[3 * i], attrib.vertices[3 * i + 1], attrib.vertices[3 * i + 2]);
        update_minmax(point, min_box, max_box);
        vertList.push_back(point);
    }
    for (size_t i = 0; i < Ftmp.size() / 3; i++) faceList.push_back(Vec3ui(Ftmp[3 * i], Ftmp[3 * i + 1], Ftmp[3 * i + 2]));
    if (is_unit_cube) { min_box = Vec3f(-1, -1, -1) * 0.5; max_box = Vec3f(1, 1, 1) * 0.5; }      // main.cpp:82-85
    Vec3f unit(1, 1, 1);
    min_box -= padding * dx * unit;                                                    // main.cpp:91-94
    max_box += padding * dx * unit;
    Vec3ui sizes = Vec3ui((max_box - min_box) / dx);

    Array3f phi_grid;
    make_level_set3(faceList, vertList, min_box, dx, sizes[0], sizes[1], sizes[2], phi_grid);
    assert(phi_grid.ni == dim && phi_grid.nj == dim && phi_grid.nk == dim);           // main.cpp:102

    int32_t dims[3] = { phi_grid.ni, phi_grid.nj, phi_grid.nk };                       // main.cpp:104-120 + save_vox
    float res = 1.0, dfix = 1.0 / dim, g2w[16] = { 0 };
    float tr = (padding != 0) ? (float)(-0.5 * 1.0f) - dfix : (float)(-0.5 * 1.0f);
    g2w[0] = g2w[5] = g2w[10] = 1.0f * dx;
    g2w[3] = g2w[7] = g2w[11] = tr;
    g2w[15] = 1.0f;
    std::vector<float> sdf(phi_grid.a.size());
    for (unsigned int i = 0; i < phi_grid.a.size(); ++i) sdf[i] = phi_grid.a[i] * dim;
    FILE* f = std::fopen(filename_out.c_str(), "wb");
    assert(f);
    std::fwrite(dims, 4, 3, f);
    std::fwrite(&res, 4, 1, f);
    std::fwrite(g2w, 4, 16, f);
    std::fwrite(sdf.data(), 4, sdf.size(), f);
    std::fclose(f);
    return 0;
}

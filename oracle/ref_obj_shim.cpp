// oracle/ref_obj_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-linkage doorway into the OBJ reader the reference uses: its vendored src/utils/tiny_obj_loader.h, compiled IN
// PLACE from /root/reference (never copied into this repo). load_mesh itself (src/utils/LoaderMesh.h:15-84) needs
// Eigen, which is not installed here; its OBJ branch (:17-50) is restated below without Eigen: every attrib.vertices
// entry, and the vertex_index of every corner of every face of every shape, in order. Built only by oracle/Makefile
// into oracle/_ref/libg3d_ref_obj.so (git-ignored; travels to the GPU box).
#define TINYOBJLOADER_IMPLEMENTATION
#include G3D_REF_TINYOBJ_H   // -D'G3D_REF_TINYOBJ_H="/root/reference/src/utils/tiny_obj_loader.h"'

#include <cstdint>
#include <cstdlib>
#include <cstring>

extern "C" {

// returns 0 on success, 1 when tinyobj::LoadObj returned false (LoaderMesh.h:24 asserts there)
int g3d_ref_load_obj(const char* path, float** verts, int64_t* n_vert, int32_t** faces, int64_t* n_index)
{
    std::string filename(path);
    tinyobj::attrib_t attrib;
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    std::string err;
    std::string materialDir = filename.substr(0, filename.rfind("/") + 1);            // LoaderMesh.h:21
    bool ret = tinyobj::LoadObj(&attrib, &shapes, &materials, &err, filename.c_str(), materialDir.c_str());
    *verts = nullptr; *faces = nullptr; *n_vert = 0; *n_index = 0;
    if (!ret) return 1;
    std::vector<int> Ftmp;                                                             // LoaderMesh.h:26-42
    for (size_t s = 0; s < shapes.size(); s++) {
        size_t index_offset = 0;
        for (size_t f = 0; f < shapes[s].mesh.num_face_vertices.size(); f++) {
            int fv = shapes[s].mesh.num_face_vertices[f];
            for (size_t v = 0; v < (size_t)fv; v++) Ftmp.push_back(shapes[s].mesh.indices[index_offset + v].vertex_index);
            index_offset += fv;
        }
    }
    *n_vert = (int64_t)attrib.vertices.size() / 3;
    *n_index = (int64_t)Ftmp.size();
    if (!attrib.vertices.empty()) {
        *verts = (float*)std::malloc(attrib.vertices.size() * sizeof(float));
        std::memcpy(*verts, attrib.vertices.data(), attrib.vertices.size() * sizeof(float));
    }
    if (!Ftmp.empty()) {
        *faces = (int32_t*)std::malloc(Ftmp.size() * sizeof(int32_t));
        std::memcpy(*faces, Ftmp.data(), Ftmp.size() * sizeof(int32_t));
    }
    return 0;
}

void g3d_ref_obj_free(void* p) { std::free(p); }

}  // extern "C"

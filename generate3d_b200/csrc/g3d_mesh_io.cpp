// g3d_mesh_io.cpp -- host-side mesh reader behind g3d_load_mesh.
//
// Stands in for load_mesh (src/utils/LoaderMesh.h:15-84): Wavefront OBJ through
// the subset of tinyobjloader's behaviour the dfgen path relies on -- every `v`
// record is kept (also unreferenced ones, main.cpp:71-76 copies all columns),
// faces are emitted in file order across groups, indices may be negative
// (relative) and may carry /vt/vn suffixes -- and ASCII / binary little-endian
// PLY with float x,y,z and a list face property. Triangle faces map 1:1.
// Polygons with more than three corners are fan-triangulated (the reference's
// tinyobj version ear-clips them; dfgen's usage text demands triangle meshes).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "g3d_internal.h"

namespace {

bool ends_with_ci(const std::string& s, const char* ext)
{
    return s.find(ext) != std::string::npos;       // LoaderMesh.h:16,53 uses find(), not a suffix test
}

int load_obj(const char* path, std::vector<float>& v, std::vector<uint32_t>& f)
{
    FILE* fp = fopen(path, "rb");
    if (!fp) { g3d::set_error("load_mesh: cannot open %s", path); return G3D_ERR_IO; }
    std::vector<char> line;
    std::vector<long long> corner;
    int c;
    bool bad = false;
    for (;;) {
        line.clear();
        while ((c = fgetc(fp)) != EOF && c != '\n') line.push_back((char)c);
        if (c == EOF && line.empty()) break;
        line.push_back('\0');
        const char* p = line.data();
        while (*p == ' ' || *p == '\t') ++p;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            char* e = nullptr;
            p += 2;
            for (int k = 0; k < 3; ++k) {
                float x = strtof(p, &e);
                if (e == p) x = 0.f;               // tinyobj's parseReal defaults to 0
                v.push_back(x);
                p = e;
            }
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            p += 2;
            corner.clear();
            const long long nv = (long long)v.size() / 3;
            for (;;) {
                while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
                if (!*p) break;
                char* e = nullptr;
                long long idx = strtoll(p, &e, 10);
                if (e == p) { bad = true; break; }
                // tinyobj fixIndex: 1-based, negative = relative to the vertices seen so far
                long long vi = idx > 0 ? idx - 1 : (idx < 0 ? nv + idx : -1);
                corner.push_back(vi);
                p = e;
                while (*p && *p != ' ' && *p != '\t' && *p != '\r') ++p;   // skip /vt/vn
            }
            for (size_t k = 2; k < corner.size(); ++k) {
                long long t[3] = { corner[0], corner[k - 1], corner[k] };
                for (int q = 0; q < 3; ++q) f.push_back(t[q] < 0 ? 0xffffffffu : (uint32_t)t[q]);
            }
        }
        if (c == EOF) break;
    }
    fclose(fp);
    if (bad) { g3d::set_error("load_mesh: malformed face record in %s", path); return G3D_ERR_IO; }
    return G3D_OK;
}

struct PlyProp { std::string type, name; bool list = false; std::string count_type; };

size_t ply_size(const std::string& t)
{
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}

double ply_read_bin(std::istream& in, const std::string& t)
{
    unsigned char b[8] = { 0 };
    size_t n = ply_size(t);
    in.read((char*)b, (std::streamsize)n);
    if (t == "float" || t == "float32") { float x; memcpy(&x, b, 4); return x; }
    if (t == "double" || t == "float64") { double x; memcpy(&x, b, 8); return x; }
    if (t == "char" || t == "int8") return (int8_t)b[0];
    if (t == "uchar" || t == "uint8") return b[0];
    if (t == "short" || t == "int16") { int16_t x; memcpy(&x, b, 2); return x; }
    if (t == "ushort" || t == "uint16") { uint16_t x; memcpy(&x, b, 2); return x; }
    if (t == "int" || t == "int32") { int32_t x; memcpy(&x, b, 4); return x; }
    uint32_t x; memcpy(&x, b, 4); return x;
}

int load_ply(const char* path, std::vector<float>& v, std::vector<uint32_t>& f)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) { g3d::set_error("load_mesh: cannot open %s", path); return G3D_ERR_IO; }
    std::string line, fmt;
    struct Elem { std::string name; long long count = 0; std::vector<PlyProp> props; };
    std::vector<Elem> elems;
    std::getline(in, line);
    if (line.compare(0, 3, "ply") != 0) { g3d::set_error("load_mesh: %s is not a PLY file", path); return G3D_ERR_IO; }
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ss(line);
        std::string tok;
        ss >> tok;
        if (tok == "format") ss >> fmt;
        else if (tok == "element") { Elem e; ss >> e.name >> e.count; elems.push_back(e); }
        else if (tok == "property" && !elems.empty()) {
            PlyProp p; ss >> p.type;
            if (p.type == "list") { p.list = true; ss >> p.count_type >> p.type; }
            ss >> p.name;
            elems.back().props.push_back(p);
        } else if (tok == "end_header") break;
    }
    const bool ascii = fmt == "ascii";
    if (!ascii && fmt != "binary_little_endian") { g3d::set_error("load_mesh: PLY format '%s' not supported", fmt.c_str()); return G3D_ERR_IO; }
    auto next = [&](const std::string& t) -> double {
        if (ascii) { double x = 0; in >> x; return x; }
        return ply_read_bin(in, t);
    };
    for (const Elem& e : elems)
        for (long long r = 0; r < e.count; ++r) {
            float xyz[3] = { 0, 0, 0 };
            for (const PlyProp& p : e.props) {
                if (p.list) {
                    long long n = (long long)next(p.count_type);
                    std::vector<long long> c((size_t)(n > 0 ? n : 0));
                    for (long long k = 0; k < n; ++k) c[(size_t)k] = (long long)next(p.type);
                    if (e.name == "face" && (p.name == "vertex_indices" || p.name == "vertex_index"))
                        for (size_t k = 2; k < c.size(); ++k) { f.push_back((uint32_t)c[0]); f.push_back((uint32_t)c[k - 1]); f.push_back((uint32_t)c[k]); }
                } else {
                    double x = next(p.type);
                    if (e.name == "vertex") {
                        if (p.name == "x") xyz[0] = (float)x;
                        else if (p.name == "y") xyz[1] = (float)x;
                        else if (p.name == "z") xyz[2] = (float)x;
                    }
                }
            }
            if (e.name == "vertex") { v.push_back(xyz[0]); v.push_back(xyz[1]); v.push_back(xyz[2]); }
            if (!in) { g3d::set_error("load_mesh: truncated PLY %s", path); return G3D_ERR_IO; }
        }
    return G3D_OK;
}

}  // namespace

extern "C" int g3d_load_mesh(const char* path, float** verts, int64_t* n_vert, uint32_t** tris, int64_t* n_tri)
{
    if (!path || !verts || !n_vert || !tris || !n_tri) { g3d::set_error("load_mesh: NULL argument"); return G3D_ERR_INVALID; }
    *verts = nullptr; *tris = nullptr; *n_vert = 0; *n_tri = 0;
    std::vector<float> v;
    std::vector<uint32_t> f;
    std::string name(path);
    int rc;
    if (ends_with_ci(name, ".obj")) rc = load_obj(path, v, f);
    else if (ends_with_ci(name, ".ply")) rc = load_ply(path, v, f);
    else { g3d::set_error("Error: Mesh format not known."); return G3D_ERR_INVALID; }   // LoaderMesh.h:79-82
    if (rc != G3D_OK) return rc;
    *n_vert = (int64_t)v.size() / 3;
    *n_tri = (int64_t)f.size() / 3;
    if (!v.empty()) { *verts = (float*)malloc(v.size() * sizeof(float)); memcpy(*verts, v.data(), v.size() * sizeof(float)); }
    if (!f.empty()) { *tris = (uint32_t*)malloc(f.size() * sizeof(uint32_t)); memcpy(*tris, f.data(), f.size() * sizeof(uint32_t)); }
    return G3D_OK;
}

extern "C" void g3d_free(void* p) { free(p); }

// g3d_mesh_io.cpp -- host-side mesh reader behind g3d_load_mesh.
//
// Stands in for load_mesh (src/utils/LoaderMesh.h:15-84): Wavefront OBJ through
// the subset of tinyobjloader's behaviour the dfgen path relies on -- every `v`
// record is kept (also unreferenced ones, main.cpp:71-76 copies all columns),
// faces are emitted in file order across groups, indices may be negative
// (relative) and may carry /vt/vn suffixes -- and ASCII / binary little-endian
// PLY with float x,y,z and a list face property. Numbers and polygons are
// handled exactly as the vendored tinyobjloader handles them (see load_obj);
// pinned against that header compiled in place (tests/test_mesh_io.py).
#include <math.h>
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

// ---- Wavefront OBJ, with the observable behaviour of the tinyobjloader version the reference vendors
// (src/utils/tiny_obj_loader.h) as load_mesh uses it (LoaderMesh.h:17-50: every attrib.vertices entry, the
// vertex_index of every emitted corner, shapes in order). What has to match for a byte-identical .df:
//   * numbers: tinyobj does not call strtod. Its parser (tiny_obj_loader.h:505-618) accumulates the integer digits
//     in a double, adds each fraction digit times a table entry 1e-k (std::pow(10, -k) from the 8th digit on), and
//     applies a decimal exponent as ldexp(m * 5^e, e); the double is then narrowed to float (:620-628). The result
//     is NOT the correctly rounded float of the text in general, so the same arithmetic is replayed here;
//     a token it cannot parse ("", ".5", "-.5", "nan") leaves the default 0;
//   * lines end at LF, CRLF or a lone CR (:399-428);
//   * face corners are read with atoi + strcspn (:754-806), so "3/ 4" and friends behave as there; index 0 (or a
//     non-number) makes the whole load fail (:1668-1676; LoaderMesh.h:24 then asserts);
//   * polygons with more than three corners are ear-clipped in the projection chosen from the first non-degenerate
//     corner (:1011-1146), float arithmetic, at most ten trips round the polygon, ears emitted in discovery order;
//     faces with fewer than three corners emit nothing;
//   * faces are flushed into the current shape at `g`, `o`, a `usemtl` that changes the material id, and at the
//     end (:1693-1911, 1939-1948). A shape whose faces were all flushed by `usemtl` before an `o` line is DROPPED
//     by that version (`o` only keeps the shape when the pending group was non-empty); reproduced, which is why the
//     material names of the mtllib files are read too (only their identity matters).
double tinyobj_number(const char* s, const char* end, double fallback)
{
    if (s >= end) return fallback;
    const char* c = s;
    double sign = 1.0;
    if (*c == '+' || *c == '-') { if (*c == '-') sign = -1.0; ++c; }
    else if (!(*c >= '0' && *c <= '9')) return fallback;
    double m = 0.0;
    int nread = 0;
    while (c != end && *c >= '0' && *c <= '9') { m *= 10; m += (int)(*c - '0'); ++c; ++nread; }
    if (nread == 0) return fallback;
    int ex = 0;
    bool at_end = (c == end);
    if (!at_end) {
        if (*c == '.') {
            static const double frac[] = { 1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001 };
            ++c;
            int k = 1;
            while (c != end && *c >= '0' && *c <= '9') {
                m += (int)(*c - '0') * (k < 8 ? frac[k] : pow(10.0, -k));
                ++k;
                ++c;
            }
            at_end = (c == end);
        } else if (*c != 'e' && *c != 'E') {
            at_end = true;                                  // anything else ends the number
        }
        if (!at_end && (*c == 'e' || *c == 'E')) {
            ++c;
            bool neg = false;
            if (c != end && (*c == '+' || *c == '-')) { neg = (*c == '-'); ++c; }
            else if (!(*c >= '0' && *c <= '9')) return fallback;       // an empty exponent is a parse failure
            int nd = 0;
            while (c != end && *c >= '0' && *c <= '9') { ex *= 10; ex += (int)(*c - '0'); ++c; ++nd; }
            if (neg) ex = -ex;
            if (nd == 0) return fallback;
        }
    }
    return sign * (ex ? ldexp(m * pow(5.0, ex), ex) : m);
}

float tinyobj_real(const char*& p)
{
    p += strspn(p, " \t");
    const char* end = p + strcspn(p, " \t\r");
    const float f = (float)tinyobj_number(p, end, 0.0);
    p = end;
    return f;
}

// 1-based -> 0-based, negative = relative to the `n` records seen so far; 0 is invalid
bool obj_index(int idx, long long n, long long& out)
{
    if (idx > 0) { out = idx - 1; return true; }
    if (idx < 0) { out = n + idx; return true; }
    return false;
}

// one face corner "i", "i/j", "i//k", "i/j/k": only the vertex index is kept, but every field must be valid
bool obj_corner(const char*& p, long long nv, long long nvn, long long nvt, long long& vi)
{
    long long dummy;
    if (!obj_index(atoi(p), nv, vi)) return false;
    p += strcspn(p, "/ \t\r");
    if (p[0] != '/') return true;
    ++p;
    if (p[0] == '/') {                                       // i//k
        ++p;
        if (!obj_index(atoi(p), nvn, dummy)) return false;
        p += strcspn(p, "/ \t\r");
        return true;
    }
    if (!obj_index(atoi(p), nvt, dummy)) return false;       // i/j or i/j/k
    p += strcspn(p, "/ \t\r");
    if (p[0] != '/') return true;
    ++p;
    if (!obj_index(atoi(p), nvn, dummy)) return false;
    p += strcspn(p, "/ \t\r");
    return true;
}

int crossing_parity(const float* vx, const float* vy, float tx, float ty)       // point in the triangle (vx, vy)?
{
    int c = 0;
    for (int i = 0, j = 2; i < 3; j = i++)
        if (((vy[i] > ty) != (vy[j] > ty)) && (tx < (vx[j] - vx[i]) * (ty - vy[i]) / (vy[j] - vy[i]) + vx[i])) c = !c;
    return c;
}

// Ear clipping of one polygon as tiny_obj_loader.h:1011-1146 does it. `v` holds the vertices read SO FAR (the
// reference triangulates at flush time); a corner outside that range is undefined behaviour there and flagged here.
void emit_face(const std::vector<long long>& face, const std::vector<float>& v, std::vector<uint32_t>& out)
{
    const size_t n = face.size();
    auto put = [&](long long a) { out.push_back((a < 0 || a > 0xfffffffell) ? 0xffffffffu : (uint32_t)a); };
    if (n < 3) return;
    if (n == 3) { put(face[0]); put(face[1]); put(face[2]); return; }
    const long long nv = (long long)v.size() / 3;
    for (long long a : face)
        if (a < 0 || a >= nv) {                              // cannot be triangulated: emit a fan of flagged corners
            for (size_t k = 2; k < n; ++k) { put(-1); put(-1); put(-1); }
            return;
        }
    auto X = [&](long long a, size_t ax) { return v[(size_t)a * 3 + ax]; };
    size_t ax0 = 1, ax1 = 2;
    for (size_t k = 0; k < n; ++k) {
        const long long a = face[k % n], b = face[(k + 1) % n], c = face[(k + 2) % n];
        const float e0x = X(b, 0) - X(a, 0), e0y = X(b, 1) - X(a, 1), e0z = X(b, 2) - X(a, 2);
        const float e1x = X(c, 0) - X(b, 0), e1y = X(c, 1) - X(b, 1), e1z = X(c, 2) - X(b, 2);
        const float cx = fabsf(e0y * e1z - e0z * e1y), cy = fabsf(e0z * e1x - e0x * e1z), cz = fabsf(e0x * e1y - e0y * e1x);
        const float eps = 0.0001f;
        if (cx > eps || cy > eps || cz > eps) {
            if (!(cx > cy && cx > cz)) {
                ax0 = 0;
                if (cz > cx && cz > cy) ax1 = 1;
            }
            break;
        }
    }
    float area = 0;
    for (size_t k = 0; k < n; ++k) {
        const long long a = face[k % n], b = face[(k + 1) % n];
        area += (X(a, ax0) * X(b, ax1) - X(a, ax1) * X(b, ax0)) * 0.5f;
    }
    std::vector<long long> rest(face);
    int rounds = 10;
    size_t guess = 0;
    while (rest.size() > 3 && rounds > 0) {
        const size_t m = rest.size();
        if (guess >= m) { --rounds; guess -= m; }
        long long ind[3];
        float vx[3], vy[3];
        for (size_t k = 0; k < 3; ++k) {
            ind[k] = rest[(guess + k) % m];
            vx[k] = X(ind[k], ax0);
            vy[k] = X(ind[k], ax1);
        }
        const float e0x = vx[1] - vx[0], e0y = vy[1] - vy[0], e1x = vx[2] - vx[1], e1y = vy[2] - vy[1];
        const float cross = e0x * e1y - e0y * e1x;
        if (cross * area < 0.0f) { ++guess; continue; }      // a reflex corner
        bool overlap = false;
        for (size_t o = 3; o < m; ++o) {
            const long long q = rest[(guess + o) % m];
            if (crossing_parity(vx, vy, X(q, ax0), X(q, ax1))) { overlap = true; break; }
        }
        if (overlap) { ++guess; continue; }
        put(ind[0]); put(ind[1]); put(ind[2]);               // an ear
        rest.erase(rest.begin() + (long)((guess + 1) % m));
    }
    if (rest.size() == 3) { put(rest[0]); put(rest[1]); put(rest[2]); }
}

// names of the materials of one .mtl file, in tinyobj's numbering (tiny_obj_loader.h:1206-1262, 1570-1574):
// a material is registered when the next `newmtl` comes (unless its name is empty) and once more at the end
void read_mtl_names(FILE* fp, std::vector<std::string>& names, int& n_materials)
{
    std::string line, cur;
    int c;
    auto reg = [&](const std::string& nm) {
        bool have = false;
        for (size_t i = 0; i < names.size(); i += 1) if (names[i] == nm) { have = true; break; }
        (void)have;
        names.push_back(nm);                                 // slot index = material id (first definition wins on lookup)
        ++n_materials;
    };
    for (;;) {
        line.clear();
        bool eof = false;
        for (;;) {
            c = fgetc(fp);
            if (c == EOF) { eof = true; break; }
            if (c == '\n') break;
            if (c == '\r') { int d = fgetc(fp); if (d != '\n' && d != EOF) ungetc(d, fp); break; }
            line.push_back((char)c);
        }
        if (eof && line.empty()) break;
        const size_t last = line.find_last_not_of(" \t");
        line = (last == std::string::npos) ? std::string() : line.substr(0, last + 1);
        const char* p = line.c_str();
        p += strspn(p, " \t");
        if (strncmp(p, "newmtl", 6) == 0 && (p[6] == ' ' || p[6] == '\t')) {
            if (!cur.empty()) reg(cur);
            cur = std::string(p + 7);
        }
        if (eof) break;
    }
    reg(cur);
}

int load_obj(const char* path, std::vector<float>& v, std::vector<uint32_t>& f)
{
    FILE* fp = fopen(path, "rb");
    if (!fp) { g3d::set_error("load_mesh: cannot open %s", path); return G3D_ERR_IO; }
    std::string dir(path);
    dir = dir.substr(0, dir.rfind('/') + 1);                 // LoaderMesh.h:21 (npos + 1 == 0: empty prefix)
    std::vector<std::string> mat_names;                      // index = material id
    int n_materials = 0, material = -1;
    long long nvn = 0, nvt = 0;
    std::vector<std::vector<long long> > group;              // faces waiting to be flushed into the shape
    std::vector<uint32_t> shape;                             // corners of the current shape
    std::string line;
    bool bad = false;
    auto flush = [&]() -> bool {
        if (group.empty()) return false;
        for (const auto& face : group) emit_face(face, v, shape);
        return true;
    };
    auto keep_shape = [&]() { f.insert(f.end(), shape.begin(), shape.end()); };
    for (bool eof = false; !eof && !bad;) {
        line.clear();
        for (;;) {
            const int c = fgetc(fp);
            if (c == EOF) { eof = true; break; }
            if (c == '\n') break;
            if (c == '\r') { const int d = fgetc(fp); if (d != '\n' && d != EOF) ungetc(d, fp); break; }
            line.push_back((char)c);
        }
        // an embedded NUL ends the line for tinyobj's C-string scanning as well (c_str())
        const char* p = line.c_str();
        p += strspn(p, " \t");
        if (p[0] == '\0' || p[0] == '#') continue;
        const bool sp1 = (p[1] == ' ' || p[1] == '\t'), sp2 = p[1] && (p[2] == ' ' || p[2] == '\t');
        if (p[0] == 'v' && sp1) {
            p += 2;
            const float x = tinyobj_real(p), y = tinyobj_real(p), z = tinyobj_real(p);
            v.push_back(x); v.push_back(y); v.push_back(z);
        } else if (p[0] == 'v' && p[1] == 'n' && sp2) {
            ++nvn;
        } else if (p[0] == 'v' && p[1] == 't' && sp2) {
            ++nvt;
        } else if (p[0] == 'f' && sp1) {
            p += 2;
            p += strspn(p, " \t");
            std::vector<long long> face;
            while (!(p[0] == '\r' || p[0] == '\n' || p[0] == '\0')) {
                long long vi;
                if (!obj_corner(p, (long long)v.size() / 3, nvn, nvt, vi)) { bad = true; break; }
                face.push_back(vi);
                p += strspn(p, " \t\r");
            }
            if (!bad) group.push_back(face);
        } else if (strncmp(p, "usemtl", 6) == 0 && (p[6] == ' ' || p[6] == '\t')) {
            const std::string name(p + 7);
            int id = -1;
            for (size_t i = 0; i < mat_names.size(); ++i) if (mat_names[i] == name) { id = (int)i; break; }
            if (id != material) {
                flush();
                group.clear();
                material = id;
            }
        } else if (strncmp(p, "mtllib", 6) == 0 && (p[6] == ' ' || p[6] == '\t')) {
            std::stringstream ss{ std::string(p + 7) };
            std::string item;
            while (std::getline(ss, item, ' ')) {              // first file that opens wins (:1713-1727)
                FILE* mf = fopen((dir + item).c_str(), "rb");
                if (!mf) continue;
                read_mtl_names(mf, mat_names, n_materials);
                fclose(mf);
                break;
            }
        } else if (p[0] == 'g' && sp1) {
            flush();
            if (!shape.empty()) keep_shape();
            shape.clear();
            group.clear();
        } else if (p[0] == 'o' && sp1) {
            if (flush()) keep_shape();                         // sic: a shape filled only by usemtl flushes is dropped
            shape.clear();
            group.clear();
        }
    }
    fclose(fp);
    if (bad) { g3d::set_error("load_mesh: failed to parse an `f' line of %s (e.g. zero value for a face index)", path); return G3D_ERR_IO; }
    if (flush() || !shape.empty()) keep_shape();
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

// ------------------------------------------------------------------------------------------ vox2mesh ----
namespace {

// Colormap.h:1301-1325 with T = float (the `0.25 * dv` style terms are double arithmetic there, kept so)
void jet_color(float v, float c[3])
{
    const float vmin = (float)-0.2, vmax = (float)1.0;
    c[0] = c[1] = c[2] = 1.0f;
    if (v < vmin) v = vmin;
    if (v > vmax) v = vmax;
    const float dv = vmax - vmin;
    if (v < (vmin + 0.25 * dv)) {
        c[0] = 0;
        c[1] = 4 * (v - vmin) / dv;
    } else if (v < (vmin + 0.5 * dv)) {
        c[0] = 0;
        c[2] = (float)(1 + 4 * (vmin + 0.25 * dv - v) / dv);
    } else if (v < (vmin + 0.75 * dv)) {
        c[0] = (float)(4 * (v - vmin - 0.5 * dv) / dv);
        c[2] = 0;
    } else {
        c[1] = (float)(1 + 4 * (vmin + 0.75 * dv - v) / dv);
        c[2] = 0;
    }
}

// "%g" is what `ostream << float` prints with the default precision of 6 in the "C" locale (tinyply.h:452,616-617)
void put_float(std::string& out, float x)
{
    char buf[32];
    snprintf(buf, sizeof buf, "%g ", (double)x);
    out += buf;
}

}  // namespace

extern "C" int g3d_vox2mesh(const char* in_df, const char* out_ply, int32_t is_unitless, int32_t redcenter,
                            const char* cmap, float trunc, int32_t write_ply)
{
    if (!in_df || !out_ply) { g3d::set_error("vox2mesh: NULL path"); return G3D_ERR_INVALID; }
    if (cmap && strcmp(cmap, "jet") != 0) { g3d::set_error("vox2mesh: colour map '%s' is not carried (only jet)", cmap); return G3D_ERR_INVALID; }
    std::string txt(out_ply);
    const size_t dot = txt.rfind('.');
    if (dot == std::string::npos) { g3d::set_error("vox2mesh: output name '%s' has no extension (the reference throws here, main.cpp:196)", out_ply); return G3D_ERR_INVALID; }
    txt.replace(dot, 4, ".txt");                               // main.cpp:195-196
    // load_vox (LoaderVOX.h:20-46): header, sdf, optional pdf
    FILE* f = fopen(in_df, "rb");
    if (!f) { g3d::set_error("vox2mesh: cannot open %s", in_df); return G3D_ERR_IO; }
    int32_t dims[3];
    float res = 0.f, g2w[16];
    bool ok = fread(dims, 4, 3, f) == 3 && fread(&res, 4, 1, f) == 1 && fread(g2w, 4, 16, f) == 16;
    if (!ok || dims[0] < 0 || dims[1] < 0 || dims[2] < 0 || dims[0] > 65536 || dims[1] > 65536 || dims[2] > 65536 ||
        (int64_t)dims[0] * dims[1] * dims[2] > ((int64_t)1 << 31)) {
        fclose(f);
        g3d::set_error("vox2mesh: %s is not a .df file", in_df);
        return G3D_ERR_IO;
    }
    const size_t n = (size_t)dims[0] * dims[1] * dims[2];
    std::vector<float> sdf(n), pdf;
    ok = fread(sdf.data(), 4, n, f) == n;
    if (ok) {
        pdf.resize(n);
        if (fread(pdf.data(), 4, n, f) != n) pdf.clear();     // no (complete) pdf block
    }
    fclose(f);
    if (!ok) { g3d::set_error("vox2mesh: short read from %s", in_df); return G3D_ERR_IO; }
    if (pdf.empty()) {                                         // main.cpp:180-192
        pdf.assign(n, 0.f);
        if (redcenter) {
            const int c = dims[0] / 2, dim = dims[0], w = 1;
            for (int i = c - w; i < c + w + 1; i++)
                for (int j = c - w; j < c + w + 1; j++)
                    for (int k = c - w; k < c + w + 1; k++) {
                        const int64_t q = (int64_t)i * dim * dim + (int64_t)j * dim + k;
                        if (q >= 0 && (size_t)q < n) pdf[(size_t)q] = 1;
                    }
        }
    }
    float vs[3] = { 1.f, 1.f, 1.f };                           // voxelsize; grid2world is row-major on disk
    if (is_unitless)
        for (int a = 0; a < 3; ++a) vs[a] = sqrtf(g2w[a] * g2w[a] + g2w[4 + a] * g2w[4 + a] + g2w[8 + a] * g2w[8 + a]);   // SE3.h:118-122
    const float thr = trunc * res;
    std::string rows, verts, faces;
    size_t nvis = 0;
    static const float bv[8][3] = { { -1, -1, 1 }, { 1, -1, 1 }, { 1, 1, 1 }, { -1, 1, 1 }, { -1, -1, -1 }, { 1, -1, -1 }, { 1, 1, -1 }, { -1, 1, -1 } };   // Box3D.h:9-19
    static const int be[12][3] = { { 0, 1, 2 }, { 2, 3, 0 }, { 1, 5, 6 }, { 6, 2, 1 }, { 7, 6, 5 }, { 5, 4, 7 }, { 4, 0, 3 }, { 3, 7, 4 },
                                   { 4, 5, 1 }, { 1, 0, 4 }, { 3, 2, 6 }, { 6, 7, 3 } };                                                                   // Box3D.h:41-59
    const float r3[3] = { 0.45f * vs[0] * res, 0.45f * vs[1] * res, 0.45f * vs[2] * res };                                                                // main.cpp:72
    char buf[160];
    for (int k = 0; k < dims[2]; k++)
        for (int j = 0; j < dims[1]; j++)
            for (int i = 0; i < dims[0]; i++) {
                const size_t q = (size_t)k * dims[1] * dims[0] + (size_t)j * dims[0] + i;
                if (!(fabsf(sdf[q]) <= thr)) continue;         // main.cpp:44,92
                float col[3];
                jet_color(pdf[q], col);
                const uint32_t ci[3] = { (uint32_t)(255.0f * col[0]), (uint32_t)(255.0f * col[1]), (uint32_t)(255.0f * col[2]) };
                snprintf(buf, sizeof buf, "%d %d %d %u %u %u\n", i, j, k, ci[0], ci[1], ci[2]);
                rows += buf;
                if (write_ply) {
                    float p[3];
                    for (int a = 0; a < 3; ++a)                // (grid2world * (i, j, k, 1)).topRows(3), main.cpp:48
                        p[a] = ((g2w[4 * a] * (float)i + g2w[4 * a + 1] * (float)j) + g2w[4 * a + 2] * (float)k) + g2w[4 * a + 3] * 1.0f;
                    for (int c = 0; c < 8; ++c) {
                        for (int a = 0; a < 3; ++a) put_float(verts, bv[c][a] * r3[a] + p[a]);
                        snprintf(buf, sizeof buf, "%u %u %u \n", (uint32_t)(uint8_t)ci[0], (uint32_t)(uint8_t)ci[1], (uint32_t)(uint8_t)ci[2]);
                        verts += buf;
                    }
                    for (int e = 0; e < 12; ++e) {
                        snprintf(buf, sizeof buf, "3 %zu %zu %zu \n", nvis * 8 + be[e][0], nvis * 8 + be[e][1], nvis * 8 + be[e][2]);
                        faces += buf;
                    }
                }
                ++nvis;
            }
    FILE* o = fopen(txt.c_str(), "w");
    if (!o) { g3d::set_error("failed to open %s", txt.c_str()); return G3D_ERR_IO; }
    ok = fwrite(rows.data(), 1, rows.size(), o) == rows.size();
    ok = (fclose(o) == 0) && ok;
    if (ok && write_ply) {
        o = fopen(out_ply, "w");
        if (!o) { g3d::set_error("failed to open %s", out_ply); return G3D_ERR_IO; }
        fprintf(o, "ply\nformat ascii 1.0\ncomment generated by tinyply 2.2\nelement vertex %zu\nproperty float x\nproperty float y\n"
                   "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nelement face %zu\n"
                   "property list uchar uint vertex_indices\nend_header\n", nvis * 8, nvis * 12);
        ok = fwrite(verts.data(), 1, verts.size(), o) == verts.size() && fwrite(faces.data(), 1, faces.size(), o) == faces.size();
        ok = (fclose(o) == 0) && ok;
    }
    if (!ok) { g3d::set_error("vox2mesh: short write"); return G3D_ERR_IO; }
    return G3D_OK;
}


// dfgen_main.cpp -- drop-in for the reference's dfgen executable (src/dfgen/main.cpp).
//
//   dfgen <mesh.obj|.ply> <dim> <padding> <is_unit_cube> <out.df> [--mode faithful|exact] [--occ-txt]
//
// Same five positional arguments, same .df bytes (faithful mode), same exit
// behaviour: wrong argument count -> usage + exit(-1) (main.cpp:19-43), unknown
// mesh extension -> exit(1) (LoaderMesh.h:79-82), grid != dim -> abort like the
// reference's assert (main.cpp:102). The optional flags are extensions, and so is
//
//   dfgen --batch <list.txt> <dim> <padding> <is_unit_cube> [--mode ..] [--occ-txt] [--chunk 512] [--threads N]
//
// (SURVEY.md section 8b): list.txt holds one "<mesh> <out.df>" pair per line. One process, one CUDA context for the
// whole list: meshes are parsed on host threads, packed into pinned buffers and sent through
// g3d_tdf_batch_host_submit / _wait chunk by chunk, the next chunk being parsed while the GPU works and the previous
// one's files being written. Every model still gets the byte-identical .df (and .txt with --occ-txt) of the
// one-model form. A model that cannot be loaded is reported and skipped (exit code 4 at the end), the others are
// unaffected -- what CADVoxelization.py:28-29 achieves by ignoring the exit code of each per-model process.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/g3d.h"

static void usage()
{
    printf("SDFGen (B200) - converts closed oriented triangle meshes into grid-based signed distance fields.\n\n");
    printf("Usage: dfgen <filename> <dim> <padding> <is_unit_cube> <filename out> [--mode faithful|exact] [--occ-txt]\n");
    printf("       dfgen --batch <list.txt> <dim> <padding> <is_unit_cube> [--mode ..] [--occ-txt] [--chunk 512] [--threads N]\n\n");
    printf("Where:\n");
    printf("\t<filename> a Wavefront OBJ (.obj) or PLY (.ply) *triangle* mesh.\n");
    printf("\t<dim> number of grid nodes per axis.\n");
    printf("\t<padding> cells of padding around the bound box.\n");
    printf("\t<is_unit_cube> 1: the grid spans [-0.5,0.5]^3, 0: the mesh bound box.\n");
    printf("\t<filename out> E.g. out.df: int32 dims[3], f32 res, f32 grid2world[16], f32 sdf[n].\n\n");
}


static int write_occ_txt(const std::string& df_path, const float* occ, int ni, int nj, int nk)
{
    // vox2mesh write_txt (vox2mesh/main.cpp:86-104), colour jet(0)
    std::string txt(df_path);
    size_t dot = txt.rfind('.');
    txt = (dot == std::string::npos ? txt : txt.substr(0, dot)) + ".txt";
    FILE* f = fopen(txt.c_str(), "w");
    if (!f) { fprintf(stderr, "failed to open %s\n", txt.c_str()); return 2; }
    for (int k = 0; k < nk; ++k) for (int j = 0; j < nj; ++j) for (int i = 0; i < ni; ++i)
        if (occ[(size_t)i + (size_t)ni * (j + (size_t)nj * k)] != 0.f) fprintf(f, "%d %d %d 0 169 255\n", i, j, k);
    fclose(f);
    return 0;
}

template <class F> static void parallel_for(int n, int threads, F fn)
{
    std::atomic<int> next(0);
    auto work = [&]() { for (int i; (i = next.fetch_add(1)) < n;) fn(i); };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads && t < n; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
}

struct Job { std::string mesh, out; float* v = nullptr; uint32_t* t = nullptr; int64_t nv = 0, nt = 0; bool ok = false; std::string err; };

struct HostSet {                                    // one set of pinned staging buffers per batch in flight
    float* v = nullptr; uint32_t* t = nullptr; float* sdf = nullptr; float* occ = nullptr; int32_t* status = nullptr;
    size_t cv = 0, ct = 0, cg = 0, co = 0, cm = 0;
    std::vector<int64_t> voff, toff;
    std::vector<int> jobs;                          // indices into the job list
    int32_t ticket = -1;
    template <class T> static bool grow(T*& p, size_t& cap, size_t need) {
        if (need <= cap) return true;
        g3d_host_free(p);
        cap = need + need / 4;
        p = (T*)g3d_host_alloc(cap * sizeof(T));
        return p != nullptr;
    }
};

static int run_batch(const char* list_path, int dim, int padding, int unit, int mode, bool occ_txt, int chunk, int threads)
{
    std::vector<Job> jobs;
    {
        FILE* lf = fopen(list_path, "r");
        if (!lf) { fprintf(stderr, "dfgen: cannot open list %s\n", list_path); return 2; }
        char a[4096], b[4096];
        while (fscanf(lf, "%4095s %4095s", a, b) == 2) { Job j; j.mesh = a; j.out = b; jobs.push_back(j); }
        fclose(lf);
    }
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (chunk < 1) chunk = 512;
    if (!unit) chunk = 1;                           // the grid follows each mesh's own bounding box (main.cpp:66-76)
    g3d_context* ctx = nullptr;
    if (g3d_context_create(0, &ctx) != G3D_OK) { fprintf(stderr, "%s\n", g3d_last_error()); return 2; }
    const int n = (int)jobs.size();
    const int64_t tri_budget = 24 << 20;            // triangles per submitted batch (~2 GB of device scratch)
    int n_fail = 0, n_done = 0;
    HostSet sets[2];
    g3d_tdf_params P[2];
    float G2W[2][16];

    auto load = [&](int lo, int hi) {
        parallel_for(hi - lo, threads, [&](int i) {
            Job& j = jobs[lo + i];
            int rc = g3d_load_mesh(j.mesh.c_str(), &j.v, &j.nv, &j.t, &j.nt);
            j.ok = rc == G3D_OK && j.nv > 0 && j.nt > 0;
            if (!j.ok) j.err = rc == G3D_OK ? "Error loading mesh." : g3d_last_error();
        });
    };
    auto release = [&](Job& j) { g3d_free(j.v); g3d_free(j.t); j.v = nullptr; j.t = nullptr; };
    // pack the loaded jobs of [lo, hi) into set s and submit; returns false when nothing was submitted
    auto submit = [&](int lo, int hi, int s) -> bool {
        HostSet& H = sets[s];
        H.jobs.clear(); H.voff.assign(1, 0); H.toff.assign(1, 0); H.ticket = -1;
        for (int i = lo; i < hi; ++i) {
            Job& j = jobs[i];
            if (!j.ok) { fprintf(stderr, "dfgen: %s: %s\n", j.mesh.c_str(), j.err.c_str()); ++n_fail; release(j); continue; }
            H.jobs.push_back(i);
            H.voff.push_back(H.voff.back() + j.nv);
            H.toff.push_back(H.toff.back() + j.nt);
        }
        const int m = (int)H.jobs.size();
        if (m == 0) return false;
        float lo3[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, hi3[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
        if (!unit) {
            const Job& j = jobs[H.jobs[0]];
            for (int64_t q = 0; q < j.nv; ++q)
                for (int a = 0; a < 3; ++a) { float x = j.v[3 * q + a]; if (x < lo3[a]) lo3[a] = x; if (x > hi3[a]) hi3[a] = x; }
        }
        memset(&P[s], 0, sizeof P[s]);
        int rc = g3d_dfgen_geometry(dim, padding, unit, lo3, hi3, &P[s], G2W[s]);
        if (rc != G3D_OK) {                         // the reference asserts (main.cpp:102): this model fails, the run goes on
            for (int i : H.jobs) { fprintf(stderr, "dfgen: %s: %s\n", jobs[i].mesh.c_str(), g3d_last_error()); ++n_fail; release(jobs[i]); }
            H.jobs.clear();
            return false;
        }
        P[s].trunc = 3.0f; P[s].occ_thresh = 1.0f; P[s].mode = mode;
        const size_t nvox = (size_t)P[s].ni * P[s].nj * P[s].nk;
        bool ok = HostSet::grow(H.v, H.cv, (size_t)H.voff.back() * 3) && HostSet::grow(H.t, H.ct, (size_t)H.toff.back() * 3) &&
                  HostSet::grow(H.sdf, H.cg, (size_t)m * nvox) && HostSet::grow(H.status, H.cm, (size_t)m);
        if (ok && occ_txt) ok = HostSet::grow(H.occ, H.co, (size_t)m * nvox);
        if (!ok) { fprintf(stderr, "dfgen: %s\n", g3d_last_error()); return false; }
        bool idx16 = true;                          // 16-bit indices when every mesh of the batch allows it: a third less PCIe traffic
        for (int q = 0; q < m; ++q) idx16 = idx16 && jobs[H.jobs[q]].nv <= 65535;
        parallel_for(m, threads, [&](int q) {
            Job& j = jobs[H.jobs[q]];
            memcpy(H.v + 3 * H.voff[q], j.v, (size_t)j.nv * 12);
            if (idx16) {
                uint16_t* dst = reinterpret_cast<uint16_t*>(H.t) + 3 * H.toff[q];
                for (int64_t e = 0; e < 3 * j.nt; ++e) dst[e] = j.t[e] > 0xffffu ? (uint16_t)0xffffu : (uint16_t)j.t[e];   // an out-of-range index stays out of range (nv <= 65535)
            } else {
                memcpy(H.t + 3 * H.toff[q], j.t, (size_t)j.nt * 12);
            }
            printf("file: %s\nn-faces: %lld\nn-verts: %lld\n", j.mesh.c_str(), (long long)j.nt, (long long)j.nv);
            release(j);
        });
        g3d_tdf_outputs out;
        memset(&out, 0, sizeof out);
        out.sdf = H.sdf; out.occ_f32 = occ_txt ? H.occ : nullptr; out.status = H.status;
        rc = idx16 ? g3d_tdf_batch_host_submit_u16(ctx, H.v, H.voff.data(), reinterpret_cast<const uint16_t*>(H.t), H.toff.data(), m, &P[s], &out, &H.ticket)
                   : g3d_tdf_batch_host_submit(ctx, H.v, H.voff.data(), H.t, H.toff.data(), m, &P[s], &out, &H.ticket);
        if (rc != G3D_OK) {
            fprintf(stderr, "dfgen: batch of %d models failed: %s\n", m, g3d_last_error());
            n_fail += m;
            H.jobs.clear();
            return false;
        }
        return true;
    };
    auto collect = [&](int s) {
        HostSet& H = sets[s];
        if (H.jobs.empty()) return;
        if (g3d_tdf_batch_host_wait(ctx, H.ticket) != G3D_OK) {
            fprintf(stderr, "dfgen: batch failed: %s\n", g3d_last_error());
            n_fail += (int)H.jobs.size();
            H.jobs.clear();
            return;
        }
        const size_t nvox = (size_t)P[s].ni * P[s].nj * P[s].nk;
        const int32_t dims[3] = { P[s].ni, P[s].nj, P[s].nk };
        std::atomic<int> bad(0);
        parallel_for((int)H.jobs.size(), threads, [&](int q) {
            const Job& j = jobs[H.jobs[q]];
            if (H.status[q]) fprintf(stderr, "warning: %s: mesh status %d (out-of-range face index or empty mesh)\n", j.mesh.c_str(), H.status[q]);
            int rc = g3d_write_df(j.out.c_str(), dims, 1.0f, G2W[s], H.sdf + (size_t)q * nvox);
            if (rc == G3D_OK && occ_txt) rc = write_occ_txt(j.out, H.occ + (size_t)q * nvox, dims[0], dims[1], dims[2]);
            if (rc != G3D_OK) { fprintf(stderr, "dfgen: %s: %s\n", j.out.c_str(), g3d_last_error()); ++bad; }
        });
        n_fail += bad.load();
        n_done += (int)H.jobs.size() - bad.load();
        H.jobs.clear();
    };

    // chunk boundaries: at most `chunk` models and tri_budget triangles (known only after loading, so a chunk is
    // loaded first and split if it came out too heavy)
    int pos = 0, k = 0;
    bool pending[2] = { false, false };
    int hi = pos + chunk < n ? pos + chunk : n;
    if (n > 0) load(pos, hi);
    while (pos < n) {
        int cut = pos;
        int64_t tris = 0;
        while (cut < hi && (cut == pos || tris + (jobs[cut].ok ? jobs[cut].nt : 0) <= tri_budget)) { tris += jobs[cut].ok ? jobs[cut].nt : 0; ++cut; }
        const int s = k & 1;
        if (pending[s]) { collect(s); pending[s] = false; }
        pending[s] = submit(pos, cut, s);
        if (pending[s ^ 1]) { collect(s ^ 1); pending[s ^ 1] = false; }      // overlaps the batch just submitted
        pos = cut;
        ++k;
        if (pos == hi && pos < n) { hi = pos + chunk < n ? pos + chunk : n; load(pos, hi); }   // parsed while the GPU works
    }
    for (int s = 0; s < 2; ++s) if (pending[s]) collect(s);
    for (int s = 0; s < 2; ++s) { g3d_host_free(sets[s].v); g3d_host_free(sets[s].t); g3d_host_free(sets[s].sdf); g3d_host_free(sets[s].occ); g3d_host_free(sets[s].status); }
    g3d_context_destroy(ctx);
    printf("dfgen --batch: %d of %d models written, %d failed\n", n_done, n, n_fail);
    return n_fail ? 4 : 0;
}

int main(int argc, char** argv)
{
    std::vector<const char*> pos;
    int mode = G3D_MODE_FAITHFUL;
    bool occ_txt = false;
    const char* batch_list = nullptr;
    int chunk = 512, threads = 0;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--mode") && i + 1 < argc) { mode = !strcmp(argv[++i], "exact") ? G3D_MODE_EXACT : G3D_MODE_FAITHFUL; }
        else if (!strcmp(argv[i], "--occ-txt")) occ_txt = true;
        else if (!strcmp(argv[i], "--batch") && i + 1 < argc) batch_list = argv[++i];
        else if (!strcmp(argv[i], "--chunk") && i + 1 < argc) chunk = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else pos.push_back(argv[i]);
    }
    if (batch_list) {
        if (pos.size() != 3) { usage(); exit(-1); }
        return run_batch(batch_list, atoi(pos[0]), atoi(pos[1]), atoi(pos[2]), mode, occ_txt, chunk, threads);
    }
    if (pos.size() != 5) { usage(); exit(-1); }
    const int dim = atoi(pos[1]), padding = atoi(pos[2]), unit = atoi(pos[3]);

    float* verts = nullptr; uint32_t* tris = nullptr; int64_t nv = 0, nt = 0;
    int rc = g3d_load_mesh(pos[0], &verts, &nv, &tris, &nt);
    if (rc == G3D_ERR_INVALID) { fprintf(stderr, "%s\n", g3d_last_error()); exit(1); }
    if (rc != G3D_OK || nv == 0 || nt == 0) { fprintf(stderr, "Error loading mesh. %s\n", g3d_last_error()); abort(); }
    printf("file: %s\nn-faces: %lld\nn-verts: %lld\n", pos[0], (long long)nt, (long long)nv);

    float lo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, hi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
    for (int64_t i = 0; i < nv; ++i)
        for (int a = 0; a < 3; ++a) { float x = verts[3 * i + a]; if (x < lo[a]) lo[a] = x; if (x > hi[a]) hi[a] = x; }
    g3d_tdf_params p;
    memset(&p, 0, sizeof p);
    float g2w[16];
    rc = g3d_dfgen_geometry(dim, padding, unit, lo, hi, &p, g2w);
    printf("Read in %lld vertices and %lld faces.\n", (long long)nv, (long long)nt);
    printf("padding: %d dx: %g\n", padding, p.dx);
    if (rc != G3D_OK) { fprintf(stderr, "%s\n", g3d_last_error()); abort(); }
    p.trunc = 3.0f; p.occ_thresh = 1.0f; p.mode = mode;

    g3d_context* ctx = nullptr;
    if (g3d_context_create(0, &ctx) != G3D_OK) { fprintf(stderr, "%s\n", g3d_last_error()); return 2; }
    const size_t n = (size_t)p.ni * p.nj * p.nk;
    std::vector<float> sdf(n), occ(occ_txt ? n : 0);
    int32_t status = 0;
    int64_t voff[2] = { 0, nv }, toff[2] = { 0, nt };
    g3d_tdf_outputs out;
    memset(&out, 0, sizeof out);
    out.sdf = sdf.data();
    out.occ_f32 = occ_txt ? occ.data() : nullptr;
    out.status = &status;
    rc = g3d_tdf_batch_host(ctx, verts, voff, tris, toff, 1, &p, &out);
    if (rc != G3D_OK) { fprintf(stderr, "%s\n", g3d_last_error()); return 2; }
    if (status) fprintf(stderr, "warning: mesh status %d (out-of-range face index or empty mesh)\n", status);
    int32_t dims[3] = { p.ni, p.nj, p.nk };
    if (g3d_write_df(pos[4], dims, 1.0f, g2w, sdf.data()) != G3D_OK) { fprintf(stderr, "%s\n", g3d_last_error()); return 2; }
    if (occ_txt && write_occ_txt(pos[4], occ.data(), p.ni, p.nj, p.nk) != 0) return 2;
    g3d_context_destroy(ctx);
    g3d_free(verts);
    g3d_free(tris);
    return 0;
}

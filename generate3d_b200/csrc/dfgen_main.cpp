// dfgen_main.cpp -- drop-in for the reference's dfgen executable (src/dfgen/main.cpp).
//
//   dfgen <mesh.obj|.ply> <dim> <padding> <is_unit_cube> <out.df> [--mode faithful|exact] [--occ-txt]
//
// Same five positional arguments, same .df bytes (faithful mode), same exit
// behaviour: wrong argument count -> usage + exit(-1) (main.cpp:19-43), unknown
// mesh extension -> exit(1) (LoaderMesh.h:79-82), grid != dim -> abort like the
// reference's assert (main.cpp:102). The two optional flags are extensions.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/g3d.h"

static void usage()
{
    printf("SDFGen (B200) - converts closed oriented triangle meshes into grid-based signed distance fields.\n\n");
    printf("Usage: dfgen <filename> <dim> <padding> <is_unit_cube> <filename out> [--mode faithful|exact] [--occ-txt]\n\n");
    printf("Where:\n");
    printf("\t<filename> a Wavefront OBJ (.obj) or PLY (.ply) *triangle* mesh.\n");
    printf("\t<dim> number of grid nodes per axis.\n");
    printf("\t<padding> cells of padding around the bound box.\n");
    printf("\t<is_unit_cube> 1: the grid spans [-0.5,0.5]^3, 0: the mesh bound box.\n");
    printf("\t<filename out> E.g. out.df: int32 dims[3], f32 res, f32 grid2world[16], f32 sdf[n].\n\n");
}

int main(int argc, char** argv)
{
    std::vector<const char*> pos;
    int mode = G3D_MODE_FAITHFUL;
    bool occ_txt = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--mode") && i + 1 < argc) { mode = !strcmp(argv[++i], "exact") ? G3D_MODE_EXACT : G3D_MODE_FAITHFUL; }
        else if (!strcmp(argv[i], "--occ-txt")) occ_txt = true;
        else pos.push_back(argv[i]);
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
    if (occ_txt) {                                  // vox2mesh write_txt (vox2mesh/main.cpp:86-104), colour jet(0)
        std::string txt(pos[4]);
        size_t dot = txt.rfind('.');
        txt = (dot == std::string::npos ? txt : txt.substr(0, dot)) + ".txt";
        FILE* f = fopen(txt.c_str(), "w");
        if (!f) { fprintf(stderr, "failed to open %s\n", txt.c_str()); return 2; }
        for (int k = 0; k < p.nk; ++k) for (int j = 0; j < p.nj; ++j) for (int i = 0; i < p.ni; ++i)
            if (occ[(size_t)i + (size_t)p.ni * (j + (size_t)p.nj * k)] != 0.f) fprintf(f, "%d %d %d 0 169 255\n", i, j, k);
        fclose(f);
    }
    g3d_context_destroy(ctx);
    g3d_free(verts);
    g3d_free(tris);
    return 0;
}

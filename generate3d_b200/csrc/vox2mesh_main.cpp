// vox2mesh_main.cpp -- drop-in for the reference's vox2mesh executable (src/vox2mesh/main.cpp), the second process
// CADVoxelization.py:33-37 spawns per model:
//
//   vox2mesh --in <model__0__.df> --out <model__0__.ply> [--is_unitless 0|1] [--redcenter 0|1] [--cmap jet] [--trunc 1.0]
//
// Same flags (main.cpp:131-162; `--flag value` and `--flag=value` as args.hxx accepts them), --in and --out required
// (exit 1 with a message otherwise, like the reference's validation error), same outputs: <out>.txt -- the GT
// occupancy rows load_occ reads -- and the cube-per-voxel ASCII PLY. All work is in g3d_vox2mesh (host only).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../include/g3d.h"

int main(int argc, char** argv)
{
    std::string in, out, cmap = "jet";
    bool have_in = false, have_out = false;
    int unitless = 0, redcenter = 0;
    float trunc = 1.0f;
    for (int i = 1; i < argc; ++i) {
        std::string a(argv[i]), val;
        bool has_val = false;
        const size_t eq = a.find('=');
        if (a.compare(0, 2, "--") == 0 && eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
        auto value = [&]() -> const char* {
            if (has_val) return val.c_str();
            if (i + 1 < argc) return argv[++i];
            fprintf(stderr, "Flag '%s' requires an argument but received none\n", a.c_str() + 2);
            exit(1);
        };
        if (a == "--in") { in = value(); have_in = true; }
        else if (a == "--out") { out = value(); have_out = true; }
        else if (a == "--is_unitless") unitless = atoi(value()) != 0;
        else if (a == "--redcenter") redcenter = atoi(value()) != 0;
        else if (a == "--cmap") cmap = value();
        else if (a == "--trunc") trunc = (float)atof(value());
        else { fprintf(stderr, "Flag could not be matched: %s\n", argv[i]); exit(1); }
    }
    if (!have_in || !have_out) {
        fprintf(stderr, "Group validation failed somewhere!\n  vox2mesh {OPTIONS}\n    --in=[bunny.vox]   vox file\n    --out=[bunny.ply]  out file\n"
                        "    --is_unitless=[false] --redcenter=[false] --cmap=[jet] --trunc=[1.0]\n");
        exit(1);
    }
    const int rc = g3d_vox2mesh(in.c_str(), out.c_str(), unitless, redcenter, cmap.c_str(), trunc, 1);
    if (rc != G3D_OK) { fprintf(stderr, "%s\n", g3d_last_error()); return rc == G3D_ERR_INVALID ? 1 : 2; }
    return 0;
}

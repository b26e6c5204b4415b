// g3d_internal.h -- declarations shared by the translation units of libg3d_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/g3d.h"

namespace g3d {

void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define G3D_CUDA_TRY(expr)                                          \
    do {                                                            \
        cudaError_t e__ = (expr);                                   \
        if (e__ != cudaSuccess) return g3d::cuda_fail(e__, #expr);  \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- TDF, both modes (g3d_tdf.cu), in four phases (see the definitions)
size_t tdf_workspace_bytes(int32_t n_mesh, int64_t total_tris, const g3d_tdf_params& p);
int tdf_begin(int64_t total_verts, int64_t total_tris, int32_t n_mesh, const g3d_tdf_params& p,
              const g3d_tdf_outputs& out, void* ws, size_t ws_bytes, cudaStream_t stream);
int tdf_stage_tris(const float* verts, const int64_t* vert_off, const void* tris, int index_bytes, const int64_t* tri_off,
                   int64_t t0, int64_t t1, int32_t m0, int32_t m1, int64_t total_tris, int32_t n_mesh, const g3d_tdf_params& p,
                   const g3d_tdf_outputs& out, void* ws, cudaStream_t stream);
int tdf_solve(const int64_t* tri_off, int64_t total_tris, int32_t n_mesh, const g3d_tdf_params& p, void* ws,
              cudaStream_t stream);
int tdf_finalize(const int64_t* vert_off, const int64_t* tri_off, int64_t total_tris, int32_t n_mesh, int32_t m0,
                 int32_t m1, const g3d_tdf_params& p, const g3d_tdf_outputs& out, void* ws, cudaStream_t stream);

int tdf_from_sdf_launch(const float* sdf, int64_t n, float trunc, float* tdf, float* tdflog,
                        cudaStream_t stream);
int unpack_occ_launch(const uint32_t* bits, int64_t n_bits, float* out, cudaStream_t stream);

// ---- raycast (g3d_raycast.cu)
int raycast_launch(const uint8_t* depth, const float* pose, int32_t n_view, int32_t H, int32_t W,
                   float focal, float cx, float cy, int32_t clamp_max, int32_t dim,
                   uint32_t* occ_bits, float* occ_f32, int64_t* index_map, cudaStream_t stream);

int sm_count();

// ---- profiling hooks (g3d_capi.cu); no-ops unless g3d_profile_enable was called on this thread
enum { ST_FILL = 0, ST_PREP, ST_INIT, ST_ORDER, ST_SWEEP, ST_FINAL, ST_EXACT, ST_RAYCAST, ST_COUNT };
constexpr int kProfMarks = 40;                      // events per call (chunked calls mark a stage several times)
void prof_begin(cudaStream_t s);                    // start of an API call: resets marks + counters
void prof_mark(int stage, cudaStream_t s);          // call right AFTER launching the kernel(s) of `stage`
unsigned long long* prof_counters();                // device u64[4] or nullptr

}  // namespace g3d

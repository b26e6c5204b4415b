/*
 * g3d.h -- C ABI of libg3d_b200.so, the B200 (sm_100a) implementation of
 * generate3D's voxel data-generation hot path.
 *
 * The reference has no FFI/plugin layer for this path: it
 * crosses process and file boundaries instead. Every entry point below cites
 * the reference interface it stands in for; INTEGRATION.md shows the binding a
 * maintainer of the reference would add (a ctypes stub for the Python scripts,
 * a 20-line main() for src/dfgen).
 *
 * Conventions
 *  - plain C, no torch / CUDA types in signatures; `cuda_stream` is a
 *    cudaStream_t passed as void* (NULL = legacy default stream);
 *  - the *_batch functions take DEVICE pointers owned by the caller, enqueue on
 *    the given stream and return without synchronising; they never allocate:
 *    scratch comes from the caller (`workspace`), sized by *_workspace_bytes;
 *  - the *_host functions take HOST pointers (pinned for best speed), copy in,
 *    run, copy out and synchronise the stream before returning; their device
 *    memory lives in a g3d_context created once;
 *  - return value: G3D_OK or a negative G3D_ERR_*; g3d_last_error() gives the
 *    text (thread-local). A bad mesh never aborts the batch: it is flagged in
 *    status[m] (the reference swallows per-model failures the same way,
 *    CADVoxelization.py:28-29, raytracing.py:184-195).
 *  - there is no CPU fallback: without a CUDA device every compute entry point
 *    fails with G3D_ERR_CUDA.
 */
#ifndef G3D_B200_H
#define G3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G3D_OK              0
#define G3D_ERR_INVALID    -1   /* bad argument */
#define G3D_ERR_CUDA       -2   /* CUDA runtime error / no device */
#define G3D_ERR_WORKSPACE  -3   /* workspace too small */
#define G3D_ERR_IO         -4   /* file could not be read / written */
#define G3D_ERR_ASSERT     -5   /* the reference would have tripped an assert */

#define G3D_MODE_FAITHFUL   0   /* band init + 16 sweeps: bit-exact vs make_level_set3 */
#define G3D_MODE_EXACT      1   /* brute-force nearest triangle within the truncation band */

/* per-mesh status bits written to g3d_tdf_outputs.status */
#define G3D_MESH_OK             0
#define G3D_MESH_BAD_INDEX      1   /* a face references a vertex >= n_vert (UB in the reference) */
#define G3D_MESH_EMPTY          2   /* no vertices or no faces (LoaderMesh.h:83 asserts) */
#define G3D_MESH_TOO_MANY_TRIS  4   /* more than 2^22 - 2 (exact mode: 2^27 - 2) triangles in ONE mesh: the surplus is ignored */

int         g3d_version(void);
const char* g3d_last_error(void);
/* Number of CUDA devices visible, or a negative G3D_ERR_CUDA. */
int         g3d_device_count(void);

/* ------------------------------------------------------------------ TDF ----
 * Stands in for make_level_set3 (src/dfgen/makelevelset3.h:12-14) called once
 * per model by src/dfgen/main.cpp:101, plus the scalar post-processing the
 * pipeline applies to its output:
 *   sdf      = phi * out_scale                  (main.cpp:117-120, out_scale = dim)
 *   occ bit  = |sdf| <= occ_thresh              (vox2mesh/main.cpp:92, trunc*res = 1)
 *   tdf      = sdf > trunc ? trunc : sdf        (Network/dataset_loader.py:137-139)
 *   tdflog   = sign(tdf) * log(|tdf| + 1)       (Network/losses.py:7-11)
 * Grids are laid out like the reference's Array3 / the .df payload: x fastest,
 * index = i + ni*(j + nj*k) (array3.h:68-78) -- i.e. a [nk][nj][ni] C array,
 * which is the [D][H][W] tensor load_df returns.
 */
typedef struct g3d_tdf_params {
    int32_t ni, nj, nk;     /* grid nodes per axis */
    float   origin[3];      /* world position of node (0,0,0) (main.cpp:91: min_box) */
    float   dx;             /* node spacing (main.cpp:57) */
    float   out_scale;      /* sdf = phi * out_scale */
    float   trunc;          /* truncation of the positive side, voxel units (3.0) */
    float   occ_thresh;     /* occupancy threshold, voxel units (1.0) */
    int32_t mode;           /* G3D_MODE_* */
    int32_t exact_band;     /* makelevelset3.h:14 default 1 */
} g3d_tdf_params;

/* Any output pointer may be NULL (skipped). All are [n_mesh] x grid. */
typedef struct g3d_tdf_outputs {
    float*    sdf;          /* f32 [n_mesh, nk, nj, ni] signed distance, voxel units */
    float*    tdf;          /* f32 same shape, positive side clamped at trunc */
    float*    tdflog;       /* f32 same shape */
    uint32_t* occ_bits;     /* u32 [n_mesh, ceil(ni*nj*nk/32)], bit (idx&31) of word idx>>5 */
    float*    occ_f32;      /* f32 [n_mesh, nk, nj, ni] 0/1 -- the tensor load_occ builds */
    int32_t*  closest_tri;  /* i32 same shape; faithful mode: makelevelset3.cpp closest_tri
                               (mesh-local index, -1 = none); exact mode: nearest triangle
                               or -1 beyond the truncation band */
    int32_t*  status;       /* i32 [n_mesh], G3D_MESH_* bits */
} g3d_tdf_outputs;

size_t g3d_tdf_workspace_bytes(int32_t n_mesh, int64_t total_verts, int64_t total_tris,
                               const g3d_tdf_params* params);

/*
 * verts     f32 [total_verts, 3]   all meshes back to back
 * vert_off  i64 [n_mesh + 1]       first vertex of each mesh
 * tris      u32 [total_tris, 3]    vertex indices LOCAL to their mesh
 * tri_off   i64 [n_mesh + 1]       first triangle of each mesh
 * total_verts/total_tris repeat vert_off[n_mesh]/tri_off[n_mesh] so that the
 * host never has to read device memory. One MESH may hold at most 2^22 - 2 = 4,194,302
 * triangles in faithful mode, 2^27 - 2 in exact mode (closest_tri shares a 32-bit word with
 * 10 bits of sweep bookkeeping; a larger mesh is
 * flagged G3D_MESH_TOO_MANY_TRIS and its surplus triangles are ignored); one CALL at
 * most 2^32 - 2 in total (G3D_ERR_INVALID beyond; split the batch).
 */
int g3d_tdf_batch(const float* verts, const int64_t* vert_off, int64_t total_verts,
                  const uint32_t* tris, const int64_t* tri_off, int64_t total_tris,
                  int32_t n_mesh, const g3d_tdf_params* params, const g3d_tdf_outputs* out,
                  void* workspace, size_t workspace_bytes, void* cuda_stream);

/* load_df's clamp + apply_log_transform on a device-resident sdf
 * (dataset_loader.py:137-139, losses.py:7-11); tdf or tdflog may be NULL. */
int g3d_tdf_from_sdf(const float* sdf, int64_t n, float trunc, float* tdf, float* tdflog,
                     void* cuda_stream);

/* bit-packed occupancy -> f32 0/1 grid (what load_occ returns, dataset_loader.py:102-117) */
int g3d_unpack_occ(const uint32_t* occ_bits, int64_t n_bits, float* occ_f32, void* cuda_stream);

/* -------------------------------------------------------------- raycast ----
 * Stands in for compute_index_mapping + raycast_depth
 * (src/scripts/raytracing.py:43-82, 122-144), one call for a batch of views.
 *   depth   u8  [n_view, H, W]   the 'L' PNG raster (renderer.py:34-40), 0 = background
 *   pose    f32 [n_view, 16]     row-major world->camera AFTER load_pose's patch
 *                                (dataset_loader.py:70-81)
 *   occ_bits  u32 [n_view, ceil(dim^3/32)]  bit lin = z*dim^2 + y*dim + x  ([z][y][x] grid;
 *                                the reference's txt rows are its x-major transpose)
 *   occ_f32   f32 [n_view, dim, dim, dim]   optional, [z][y][x]
 *   index_map i64 [n_view, 4, dim^3]        optional, rows umin, vmin, umax, vmax -- the
 *                                tensor compute_index_mapping returns (flattened there)
 * focal/cx/cy: make_intrinsic (raytracing.py:27-33); clamp_max: the literal 511
 * of raytracing.py:75 (kept independent of H, W exactly as in the reference;
 * slices are then clipped to the image like numpy does).
 */
int g3d_raycast_batch(const uint8_t* depth, const float* pose, int32_t n_view, int32_t H,
                      int32_t W, float focal, float cx, float cy, int32_t clamp_max, int32_t dim,
                      uint32_t* occ_bits, float* occ_f32, int64_t* index_map, void* cuda_stream);

/* ------------------------------------------------- host-buffer front end ----
 * What a cgo/ctypes/CLI caller with host arrays uses. The context owns grow-only
 * device arenas and its own streams (kernels, copies in, copies out) on `device`. */
typedef struct g3d_context g3d_context;

int  g3d_context_create(int32_t device, g3d_context** ctx);
void g3d_context_destroy(g3d_context* ctx);
/* Device bytes currently held by the context's arena. */
size_t g3d_context_arena_bytes(const g3d_context* ctx);

/* Host twin of g3d_tdf_batch: all pointers (inputs and out->*) are HOST memory. */
int g3d_tdf_batch_host(g3d_context* ctx, const float* verts, const int64_t* vert_off,
                       const uint32_t* tris, const int64_t* tri_off, int32_t n_mesh,
                       const g3d_tdf_params* params, const g3d_tdf_outputs* out);

/* The same call split in two, so that a loop over batches keeps the GPU busy: submit batch k+1, then wait for
 * batch k. Up to two batches are in flight per context (each in its own device arena); the inputs of the later one
 * cross PCIe while the earlier one is being swept, and the earlier one's results leave while the later one is
 * staged. `*ticket` identifies the batch for the wait. All host arrays of a submitted batch (inputs, offsets,
 * outputs) must stay valid and untouched until its wait returns; pinned memory makes the copies asynchronous,
 * pageable memory works but serialises. g3d_tdf_batch_host is submit followed by wait. */
int g3d_tdf_batch_host_submit(g3d_context* ctx, const float* verts, const int64_t* vert_off,
                              const uint32_t* tris, const int64_t* tri_off, int32_t n_mesh,
                              const g3d_tdf_params* params, const g3d_tdf_outputs* out, int32_t* ticket);
int g3d_tdf_batch_host_wait(g3d_context* ctx, int32_t ticket);

/* g3d_tdf_batch_host_submit with 16-bit triangle indices, for batches in which every mesh has at most 65,536 vertices
 * (G3D_ERR_INVALID otherwise): the index array is two thirds of a typical mesh's bytes, so this cuts the host->device
 * traffic of the call by about a third. Same ticket / wait protocol; tris is u16 [total_tris, 3]. */
int g3d_tdf_batch_host_submit_u16(g3d_context* ctx, const float* verts, const int64_t* vert_off,
                                  const uint16_t* tris, const int64_t* tri_off, int32_t n_mesh,
                                  const g3d_tdf_params* params, const g3d_tdf_outputs* out, int32_t* ticket);

/* Host twin of g3d_raycast_batch. */
int g3d_raycast_batch_host(g3d_context* ctx, const uint8_t* depth, const float* pose,
                           int32_t n_view, int32_t H, int32_t W, float focal, float cx, float cy,
                           int32_t clamp_max, int32_t dim, uint32_t* occ_bits, float* occ_f32,
                           int64_t* index_map);

/* Page-locked host memory for the *_host entry points (pageable memory works but serialises the copies).
 * NULL on failure (g3d_last_error has the reason). */
void* g3d_host_alloc(size_t bytes);
void  g3d_host_free(void* p);

/* ------------------------------------------------------- dfgen driver ------
 * src/dfgen/main.cpp:45-114 without the mesh: fills params (ni..dx, out_scale
 * = dim) and the row-major grid2world of the .df header. bbox_* are only read
 * when is_unit_cube == 0. Returns G3D_ERR_ASSERT when the reference would trip
 * assert(phi_grid dims == dim) (main.cpp:102). */
int g3d_dfgen_geometry(int32_t dim, int32_t padding, int32_t is_unit_cube, const float* bbox_min,
                       const float* bbox_max, g3d_tdf_params* params, float* grid2world16);

/* .df writer / reader (src/utils/LoaderVOX.h:20-63): int32 dims[3], f32 res,
 * f32 grid2world[16] row-major, f32 sdf[n] (x fastest). Host memory.
 * g3d_read_df: pass sdf == NULL to query dims only; capacity in floats. */
int g3d_write_df(const char* path, const int32_t* dims3, float res, const float* grid2world16,
                 const float* sdf);
int g3d_read_df(const char* path, int32_t* dims3, float* res, float* grid2world16, float* sdf,
                int64_t capacity);

/* vox2mesh (src/vox2mesh/main.cpp:131-203), the second process CADVoxelization.py:33-37 spawns per model, as
 * a function: reads a .df (LoaderVOX.h:20-46, optional trailing pdf block included), writes
 *   <out with its extension replaced by .txt>  rows "i j k R G B" for every voxel with |sdf| <= trunc * res in
 *                                             k, j, i loop order (main.cpp:86-104) -- the GT occupancy file
 *                                             load_occ reads; colour = cmap(pdf), jet(0) = 0 169 255;
 *   <out>                                      (when write_ply != 0) the ASCII PLY of one 0.9-voxel cube per such
 *                                             voxel that tinyply writes there (main.cpp:37-84,106-129).
 * Host-only: no GPU involved. cmap: "jet" (the default; NULL = "jet"); the other tables of Colormap.h are not
 * carried (G3D_ERR_INVALID). is_unitless != 0 scales the cubes by the column norms of grid2world (main.cpp:172-178). */
int g3d_vox2mesh(const char* in_df, const char* out_ply, int32_t is_unitless, int32_t redcenter, const char* cmap,
                 float trunc, int32_t write_ply);

/* ---------------------------------------------------- "next" rows ----
 * occ_to_df (Network/data_utils.py:115-168; occ_to_df.py:16-62): occupancy -> 26-neighbour chamfer
 * distance field, for a batch. occ f32 [n, dim, dim, dim] in [z][y][x] order; pred != 0: a cell is
 * occupied when sigmoid(x) >= 0.5 (preprocess_occ, data_utils.py:100-112), else when x == 1. The reference
 * relaxes the field exactly twice (its convergence loop ends after one trip because `df = new_df`
 * aliases the tensors), or not at all for a grid in which no free cell touches an occupied one.
 * Values above trunc become `over` (trunc in data_utils.py:164-165, trunc + 1 in occ_to_df.py:59-60).
 * out f32 [n, dim, dim, dim], same order. Device pointers. */
int g3d_occ_to_df_batch(const float* occ, int32_t n, int32_t dim, int32_t pred, float trunc, float over,
                        float* out, void* cuda_stream);

/* Colour raycast (raytracing.py:85-120): per voxel the mean colour of the pixels of its box whose depth
 * is > 0, ceil'ed; 256 = no such pixel. color u8 [n_view, H, W, 3]; out u16 [n_view, dim, dim, dim, 3]
 * in [x][y][z][rgb] order (the order in which the reference's txt rows are indexed). Device pointers. */
int g3d_raycast_color_batch(const uint8_t* depth, const uint8_t* color, const float* pose, int32_t n_view,
                            int32_t H, int32_t W, float focal, float cx, float cy, int32_t clamp_max,
                            int32_t dim, uint16_t* out, void* cuda_stream);

/* Per-pixel ray march (raytracing_approach1.py:13-51, the approach raytracing.py superseded): rays start on
 * the z_near plane and cross a camera-aligned box of side |z_far - z_near| in half-voxel steps, in float64
 * like the reference; a voxel keeps the colour of the last pixel (u outer, v inner) that touched it.
 * depth u8 [H,W], color u8 [H,W,3]; occ u8 [dim^3] and rgb u8 [dim^3,3] in [x][y][z] order (VoxelGrid's
 * own indexing); scratch u32 [dim^3]. Device pointers. */
int g3d_raymarch(const uint8_t* depth, const uint8_t* color, int32_t H, int32_t W, double cx, double cy,
                 double focal, double z_near, double z_far, int32_t dim, uint8_t* occ, uint8_t* rgb,
                 uint32_t* scratch, void* cuda_stream);

/* Z-buffered voxel -> pixel index map of the projection layer
 * (Network/perspective_projection.py:65-128, ProjectionHelper.compute_index_mapping): for every pixel the linear
 * index lin = x + dim*(y + dim*z) of the occupied voxel nearest to the camera among those whose projected box
 * covers it (depth = min |z_cam| over the 8 corners, ties to the lower index), -1 where there is none.
 * occ_bits u32 [n_grid, ceil(dim^3/32)] (bit lin&31 of word lin>>5; the reference's sigmoid(occ) > 0.2 mask),
 * pose f32 [n_grid, n_view, 16], index_map i64 [n_grid, n_view, H*W]. Device pointers; no workspace (the
 * output doubles as the key buffer). */
int g3d_zbuffer_index_map(const uint32_t* occ_bits, const float* pose, int32_t n_grid, int32_t n_view,
                          int32_t H, int32_t W, float focal, float cx, float cy, int32_t clamp_max,
                          int32_t dim, int64_t* index_map, void* cuda_stream);

/* ------------------------------------------------------------ profiling ----
 * Off by default. flags bit 0: record a CUDA event on the caller's stream between the kernels
 * of each g3d_tdf_batch / g3d_raycast_batch call (bench.py's per-kernel timing);
 * bit 1: count point-triangle evaluations with device atomics (the "instrumented run" the
 * algorithmic-FLOP figure comes from; slows the kernels, never on in a timed region).
 * g3d_profile_read waits for the last call's events. stage_ms[8]: 0 fill, 1 prep, 2 band init,
 * 3 sweep-order table, 4 sweeps, 5 finalize, 6 exact-mode main kernel, 7 raycast (0 if not run).
 * evals[4]: band-init evaluations, sweep evaluations, exact-mode evaluated pairs, exact-mode
 * culled-box tests. State is per host thread. */
#define G3D_PROF_EVENTS 1
#define G3D_PROF_COUNT  2
int g3d_profile_enable(int32_t flags);
int g3d_profile_read(float* stage_ms8, uint64_t* evals4);
/* Dependent-FFMA microbenchmark on the current device: achieved FP32 TFLOP/s (FMA = 2 flop). */
int g3d_probe_fp32_tflops(float* tflops, void* cuda_stream);

/* Mesh reader standing in for load_mesh (src/utils/LoaderMesh.h:15-84): .obj / .ply ->
 * verts f32 [n_vert,3] (ALL vertex records, referenced or not) and tris u32 [n_tri,3] in file
 * order. Buffers are malloc'ed; release with g3d_free. Unknown extension -> G3D_ERR_INVALID
 * with the reference's message (LoaderMesh.h:79-82). */
int  g3d_load_mesh(const char* path, float** verts, int64_t* n_vert, uint32_t** tris, int64_t* n_tri);
void g3d_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* G3D_B200_H */

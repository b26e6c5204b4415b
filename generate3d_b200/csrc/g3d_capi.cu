// g3d_capi.cu -- the extern "C" surface declared in include/g3d.h.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include <stdlib.h>

#include "g3d_internal.h"

namespace g3d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what)
{
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return G3D_ERR_CUDA;
}

int sm_count()
{
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        cached_dev = dev;
    }
    return cached;
}

struct Prof {
    int flags = 0;
    cudaEvent_t ev[kProfMarks] = {};
    bool have_events = false;
    int order[kProfMarks] = {};                     // stage recorded at slot i (slot 0 = begin)
    int n_marks = 0;
    unsigned long long* counters = nullptr;
};
static thread_local Prof g_prof;


void prof_begin(cudaStream_t s)
{
    Prof& p = g_prof;
    if (!p.flags) return;
    if ((p.flags & G3D_PROF_COUNT) && p.counters) cudaMemsetAsync(p.counters, 0, 4 * sizeof(unsigned long long), s);
    if (p.flags & G3D_PROF_EVENTS) {
        p.n_marks = 0;
        cudaEventRecord(p.ev[0], s);
        p.order[0] = -1;
        p.n_marks = 1;
    }
}

void prof_mark(int stage, cudaStream_t s)
{
    Prof& p = g_prof;
    if (!(p.flags & G3D_PROF_EVENTS) || p.n_marks == 0 || p.n_marks >= kProfMarks) return;
    cudaEventRecord(p.ev[p.n_marks], s);
    p.order[p.n_marks] = stage;
    ++p.n_marks;
}

unsigned long long* prof_counters() { return (g_prof.flags & G3D_PROF_COUNT) ? g_prof.counters : nullptr; }

}  // namespace g3d

using namespace g3d;

static __global__ void fp32_probe_kernel(float* out, int iters)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f,
          a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456f) out[0] = r;                  // keeps the chain alive, never true in practice
}

constexpr int kHostChunks = 4;

struct g3d_context {
    int device = 0;
    cudaStream_t stream = nullptr;                  // kernels
    cudaStream_t s_in = nullptr;                    // host->device copies of the host-buffer front end
    cudaStream_t s_out = nullptr;                   // device->host copies: a queue of their own, so that batch k's results
                                                    // leave while batch k+1's inputs arrive (pipelined submit / wait)
    // two batches can be in flight: slot b has its own arena and events
    cudaEvent_t ev_in[2][kHostChunks] = {};         // chunk landed on the device
    cudaEvent_t ev_fin[2][kHostChunks] = {};        // chunk finalised, results may leave
    cudaEvent_t ev_done[2] = {};                    // all results of the slot's batch are on the host
    bool busy[2] = { false, false };
    // did the kernel stream wait for inputs? Timed events around the kernels of the last four submissions: a batch whose
    // first kernel could only start some time after the previous batch's last one had ended found the GPU idle, i.e. the
    // host side (PCIe, host memory shared with other ranks) is the slower party. Pipelined submissions are then cut into
    // chunks, whose kernels can start on the first chunk; otherwise they go whole (fuller launches).
    cudaEvent_t ev_kbeg[4] = {}, ev_kend[4] = {};
    unsigned seq = 0;                               // submissions so far
    unsigned seq_of[2] = { 0, 0 };                  // submission number of the batch in each slot
    bool starved = false;
    char* arena[2] = { nullptr, nullptr };
    size_t arena_bytes[2] = { 0, 0 };
};

namespace {

// Grow-only device arena per slot, from the stream-ordered allocator: growing a slot's arena neither synchronises
// the device nor disturbs the other slot's batch in flight (cudaFree would do both). The slot itself is idle when
// it grows (its previous batch has been waited for), so its old block is released in order on the kernel stream
// and the new one is allocated on the input stream, which every later use of the slot is ordered after.
// Returns nullptr on failure (error already set).
char* arena_reserve(g3d_context* c, size_t bytes, int slot = 0)
{
    if (bytes <= c->arena_bytes[slot]) return c->arena[slot];
    if (c->arena[slot]) {
        cudaFreeAsync(c->arena[slot], c->stream);
        c->arena[slot] = nullptr;
        c->arena_bytes[slot] = 0;
    }
    size_t want = align_up(bytes + bytes / 8, (size_t)1 << 20);
    cudaError_t e = cudaMallocAsync((void**)&c->arena[slot], want, c->s_in);
    if (e != cudaSuccess) { c->arena[slot] = nullptr; cuda_fail(e, "cudaMallocAsync(arena)"); return nullptr; }
    c->arena_bytes[slot] = want;
    return c->arena[slot];
}

struct Carver {
    char* base; size_t off = 0;
    explicit Carver(char* b) : base(b) {}
    template <class T> T* take(size_t count) {
        size_t o = off; off = align_up(off + count * sizeof(T), 256);
        return base ? reinterpret_cast<T*>(base + o) : nullptr;
    }
};

}  // namespace

extern "C" {

int g3d_version(void) { return 100; }

const char* g3d_last_error(void) { return g_err; }

int g3d_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    return n;
}

size_t g3d_tdf_workspace_bytes(int32_t n_mesh, int64_t total_verts, int64_t total_tris, const g3d_tdf_params* params)
{
    (void)total_verts;
    if (!params || n_mesh < 0 || total_tris < 0) return 0;
    return tdf_workspace_bytes(n_mesh, total_tris, *params);
}

int g3d_tdf_batch(const float* verts, const int64_t* vert_off, int64_t total_verts, const uint32_t* tris,
                  const int64_t* tri_off, int64_t total_tris, int32_t n_mesh, const g3d_tdf_params* params,
                  const g3d_tdf_outputs* out, void* workspace, size_t workspace_bytes, void* cuda_stream)
{
    if (!params || !out) { set_error("tdf: params/out is NULL"); return G3D_ERR_INVALID; }
    if (n_mesh > 0 && (!vert_off || !tri_off)) { set_error("tdf: offsets are NULL"); return G3D_ERR_INVALID; }
    if ((total_verts > 0 && !verts) || (total_tris > 0 && !tris)) { set_error("tdf: verts/tris are NULL"); return G3D_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)cuda_stream;
    int rc = tdf_begin(total_verts, total_tris, n_mesh, *params, *out, workspace, workspace_bytes, s);
    if (rc != G3D_OK) return rc;
    rc = tdf_stage_tris(verts, vert_off, tris, 4, tri_off, 0, total_tris, 0, n_mesh, total_tris, n_mesh, *params, *out, workspace, s);
    if (rc != G3D_OK) return rc;
    rc = tdf_solve(tri_off, total_tris, n_mesh, *params, workspace, s);
    if (rc != G3D_OK) return rc;
    return tdf_finalize(vert_off, tri_off, total_tris, n_mesh, 0, n_mesh, *params, *out, workspace, s);
}

int g3d_tdf_from_sdf(const float* sdf, int64_t n, float trunc, float* tdf, float* tdflog, void* cuda_stream)
{
    return tdf_from_sdf_launch(sdf, n, trunc, tdf, tdflog, (cudaStream_t)cuda_stream);
}

int g3d_unpack_occ(const uint32_t* occ_bits, int64_t n_bits, float* occ_f32, void* cuda_stream)
{
    return unpack_occ_launch(occ_bits, n_bits, occ_f32, (cudaStream_t)cuda_stream);
}

int g3d_raycast_batch(const uint8_t* depth, const float* pose, int32_t n_view, int32_t H, int32_t W, float focal,
                      float cx, float cy, int32_t clamp_max, int32_t dim, uint32_t* occ_bits, float* occ_f32,
                      int64_t* index_map, void* cuda_stream)
{
    return raycast_launch(depth, pose, n_view, H, W, focal, cx, cy, clamp_max, dim, occ_bits, occ_f32, index_map,
                          (cudaStream_t)cuda_stream);
}

// -------------------------------------------------------------- profiling ----
int g3d_profile_enable(int32_t flags)
{
    auto& p = g3d::g_prof;
    if ((flags & G3D_PROF_EVENTS) && !p.have_events) {
        for (int i = 0; i < kProfMarks; ++i) G3D_CUDA_TRY(cudaEventCreate(&p.ev[i]));
        p.have_events = true;
    }
    if ((flags & G3D_PROF_COUNT) && !p.counters) {
        G3D_CUDA_TRY(cudaMalloc((void**)&p.counters, 4 * sizeof(unsigned long long)));
        G3D_CUDA_TRY(cudaMemset(p.counters, 0, 4 * sizeof(unsigned long long)));
    }
    p.flags = flags;
    p.n_marks = 0;
    return G3D_OK;
}

int g3d_profile_read(float* stage_ms, uint64_t* evals)
{
    auto& p = g3d::g_prof;
    if (stage_ms) {
        for (int i = 0; i < ST_COUNT; ++i) stage_ms[i] = 0.f;
        if ((p.flags & G3D_PROF_EVENTS) && p.n_marks > 1) {
            G3D_CUDA_TRY(cudaEventSynchronize(p.ev[p.n_marks - 1]));
            for (int i = 1; i < p.n_marks; ++i) {
                float ms = 0.f;
                G3D_CUDA_TRY(cudaEventElapsedTime(&ms, p.ev[i - 1], p.ev[i]));
                stage_ms[p.order[i]] += ms;
            }
        }
    }
    if (evals) {
        for (int i = 0; i < 4; ++i) evals[i] = 0;
        if ((p.flags & G3D_PROF_COUNT) && p.counters)
            G3D_CUDA_TRY(cudaMemcpy(evals, p.counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    return G3D_OK;
}

int g3d_probe_fp32_tflops(float* tflops, void* cuda_stream)
{
    if (!tflops) { set_error("probe: NULL output"); return G3D_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)cuda_stream;
    float* d = nullptr;
    G3D_CUDA_TRY(cudaMalloc((void**)&d, 4));
    cudaEvent_t e0, e1;
    G3D_CUDA_TRY(cudaEventCreate(&e0));
    G3D_CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 1 << 15, blocks = sm_count() * 8, threads = 512;
    fp32_probe_kernel<<<blocks, threads, 0, s>>>(d, 1024);                 // warm-up
    float best = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, s);
        fp32_probe_kernel<<<blocks, threads, 0, s>>>(d, iters);
        cudaEventRecord(e1, s);
        G3D_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        float tf = (float)((double)blocks * threads * iters * 8.0 * 2.0 / (ms * 1e-3) / 1e12);
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return G3D_OK;
}

// ---------------------------------------------------------------- context ----
int g3d_context_create(int32_t device, g3d_context** ctx)
{
    if (!ctx) { set_error("ctx is NULL"); return G3D_ERR_INVALID; }
    *ctx = nullptr;
    int n = 0;
    G3D_CUDA_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return G3D_ERR_INVALID; }
    G3D_CUDA_TRY(cudaSetDevice(device));
    g3d_context* c = new g3d_context();
    c->device = device;
    {   // blocks the arenas hand back stay cached in the device's pool instead of returning to the driver at every sync
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
    for (int b = 0; b < 2; ++b) {
        for (int i = 0; i < kHostChunks && e == cudaSuccess; ++i) {
            e = cudaEventCreateWithFlags(&c->ev_in[b][i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fin[b][i], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_done[b], cudaEventDisableTiming);
    }
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) {
        e = cudaEventCreate(&c->ev_kbeg[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&c->ev_kend[i]);
    }
    if (e != cudaSuccess) { g3d_context_destroy(c); return cuda_fail(e, "cudaStreamCreate/cudaEventCreate"); }
    *ctx = c;
    return G3D_OK;
}

void g3d_context_destroy(g3d_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->s_in) { cudaStreamSynchronize(c->s_in); cudaStreamDestroy(c->s_in); }
    if (c->s_out) { cudaStreamSynchronize(c->s_out); cudaStreamDestroy(c->s_out); }
    cudaDeviceSynchronize();
    for (int b = 0; b < 2; ++b) {
        for (int i = 0; i < kHostChunks; ++i) {
            if (c->ev_in[b][i]) cudaEventDestroy(c->ev_in[b][i]);
            if (c->ev_fin[b][i]) cudaEventDestroy(c->ev_fin[b][i]);
        }
        if (c->ev_done[b]) cudaEventDestroy(c->ev_done[b]);
        if (c->arena[b]) cudaFreeAsync(c->arena[b], nullptr);
    }
    for (int i = 0; i < 4; ++i) {
        if (c->ev_kbeg[i]) cudaEventDestroy(c->ev_kbeg[i]);
        if (c->ev_kend[i]) cudaEventDestroy(c->ev_kend[i]);
    }
    delete c;
}

size_t g3d_context_arena_bytes(const g3d_context* c) { return c ? c->arena_bytes[0] + c->arena_bytes[1] : 0; }

static int tdf_host_submit(g3d_context* c, const float* verts, const int64_t* vert_off, const void* tris, int index_bytes,
                           const int64_t* tri_off, int32_t n_mesh, const g3d_tdf_params* params,
                           const g3d_tdf_outputs* out, int32_t* ticket)
{
    if (!c || !params || !out || !ticket) { set_error("tdf_host: NULL argument"); return G3D_ERR_INVALID; }
    *ticket = -1;
    if (n_mesh < 0) { set_error("tdf_host: n_mesh < 0"); return G3D_ERR_INVALID; }
    if (n_mesh > 0 && (!vert_off || !tri_off)) { set_error("tdf_host: offsets are NULL"); return G3D_ERR_INVALID; }
    const int64_t nv = n_mesh ? vert_off[n_mesh] : 0, nt = n_mesh ? tri_off[n_mesh] : 0;
    if (nv < 0 || nt < 0 || (nv > 0 && !verts) || (nt > 0 && !tris)) { set_error("tdf_host: bad mesh arrays"); return G3D_ERR_INVALID; }
    for (int m = 0; m < n_mesh; ++m)
        if (vert_off[m + 1] < vert_off[m] || tri_off[m + 1] < tri_off[m]) { set_error("tdf_host: offsets not monotone"); return G3D_ERR_INVALID; }
    const int slot = !c->busy[0] ? 0 : (!c->busy[1] ? 1 : -1);
    if (slot < 0) { set_error("tdf_host: two batches are already in flight; wait for one first"); return G3D_ERR_INVALID; }
    G3D_CUDA_TRY(cudaSetDevice(c->device));
    if (n_mesh == 0) {                              // nothing to do; the ticket is still valid
        G3D_CUDA_TRY(cudaEventRecord(c->ev_done[slot], c->s_out));
        c->busy[slot] = true;
        *ticket = slot;
        return G3D_OK;
    }
    const int64_t nvox = (int64_t)params->ni * params->nj * params->nk, pw = (nvox + 31) / 32;
    const size_t ws_bytes = g3d_tdf_workspace_bytes(n_mesh, nv, nt, params);

    auto layout = [&](char* base, g3d_tdf_outputs& d, float*& dv, int64_t*& dvo, char*& dt, int64_t*& dto, void*& ws) {
        Carver k(base);
        dv = k.take<float>((size_t)nv * 3);
        dvo = k.take<int64_t>((size_t)n_mesh + 1);
        dt = k.take<char>((size_t)nt * 3 * index_bytes);
        dto = k.take<int64_t>((size_t)n_mesh + 1);
        d.sdf = out->sdf ? k.take<float>((size_t)n_mesh * nvox) : nullptr;
        d.tdf = out->tdf ? k.take<float>((size_t)n_mesh * nvox) : nullptr;
        d.tdflog = out->tdflog ? k.take<float>((size_t)n_mesh * nvox) : nullptr;
        d.occ_bits = out->occ_bits ? k.take<uint32_t>((size_t)n_mesh * pw) : nullptr;
        d.occ_f32 = out->occ_f32 ? k.take<float>((size_t)n_mesh * nvox) : nullptr;
        d.closest_tri = out->closest_tri ? k.take<int32_t>((size_t)n_mesh * nvox) : nullptr;
        d.status = out->status ? k.take<int32_t>((size_t)n_mesh) : nullptr;
        ws = k.take<char>(ws_bytes);
        return k.off;
    };
    g3d_tdf_outputs d{};
    float* dv; int64_t* dvo; char* dt; int64_t* dto; void* ws;
    size_t need = layout(nullptr, d, dv, dvo, dt, dto, ws);
    char* base = arena_reserve(c, need, slot);
    if (!base) return G3D_ERR_CUDA;
    layout(base, d, dv, dvo, dt, dto, ws);

    // The meshes cross PCIe in up to 4 chunks on the input stream; the per-triangle stage of a chunk (records,
    // boxes, parity, band init) starts as soon as that chunk has landed, so most of the host->device time hides
    // behind kernels. The sweeps then run once on the full batch; the epilogue runs per chunk and each chunk's
    // results start their way back on the output stream while the next chunk is still being finalised. With a
    // second batch submitted before the first is waited for, its inputs arrive during the first one's sweeps and
    // the first one's results leave during its staging: the kernel stream never waits for PCIe.
    cudaStream_t s = c->stream;
    cudaEvent_t* ev_in = c->ev_in[slot];
    cudaEvent_t* ev_fin = c->ev_fin[slot];
    // from here on work is enqueued: every failure path drains the three streams before returning, so that no copy is
    // still running into the caller's arrays (or the arena) after the call has reported an error
    auto bail = [&](int code) { cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(s); cudaStreamSynchronize(c->s_out); return code; };
#define G3D_TRY_BAIL(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return bail(g3d::cuda_fail(e__, #expr)); } while (0)
    // (with the other slot's batch still in flight this batch's copies hide behind that one's kernels anyway: no chunks then,
    // whole-batch launches fill the GPU better -- 4 chunks cost the 1,024-mesh step 1.5 ms in partial waves)
    const char* e_ch = getenv("G3D_HOST_CHUNKS");               // tuning aid: 1 = always in chunks, 0 = never
    const bool chunked = e_ch ? atoi(e_ch) != 0 : (!c->busy[slot ^ 1] || c->starved);
    const int nch = (n_mesh >= 64 && chunked) ? kHostChunks : 1;
    auto chunk_lo = [&](int ch) { return (int)((int64_t)n_mesh * ch / nch); };
    G3D_TRY_BAIL(cudaMemcpyAsync(dvo, vert_off, (size_t)(n_mesh + 1) * 8, cudaMemcpyHostToDevice, c->s_in));
    G3D_TRY_BAIL(cudaMemcpyAsync(dto, tri_off, (size_t)(n_mesh + 1) * 8, cudaMemcpyHostToDevice, c->s_in));
    for (int ch = 0; ch < nch; ++ch) {
        const int m0 = chunk_lo(ch), m1 = chunk_lo(ch + 1);
        const int64_t v0 = vert_off[m0], v1 = vert_off[m1], t0 = tri_off[m0], t1 = tri_off[m1];
        if (v1 > v0) G3D_TRY_BAIL(cudaMemcpyAsync(dv + 3 * v0, verts + 3 * v0, (size_t)(v1 - v0) * 12, cudaMemcpyHostToDevice, c->s_in));
        if (t1 > t0) G3D_TRY_BAIL(cudaMemcpyAsync(dt + (size_t)3 * t0 * index_bytes, (const char*)tris + (size_t)3 * t0 * index_bytes,
                                                  (size_t)(t1 - t0) * 3 * index_bytes, cudaMemcpyHostToDevice, c->s_in));
        G3D_TRY_BAIL(cudaEventRecord(ev_in[ch], c->s_in));
    }
    G3D_TRY_BAIL(cudaStreamWaitEvent(s, ev_in[0], 0));           // the offset arrays precede chunk 0 on the input stream
    const unsigned seq = c->seq++;
    c->seq_of[slot] = seq;
    G3D_TRY_BAIL(cudaEventRecord(c->ev_kbeg[seq & 3], s));       // the kernel stream has its first inputs (and is done with the batch before)
    int rc = tdf_begin(nv, nt, n_mesh, *params, d, ws, ws_bytes, s);
    if (rc != G3D_OK) return bail(rc);
    for (int ch = 0; ch < nch; ++ch) {
        const int m0 = chunk_lo(ch), m1 = chunk_lo(ch + 1);
        G3D_TRY_BAIL(cudaStreamWaitEvent(s, ev_in[ch], 0));
        rc = tdf_stage_tris(dv, dvo, dt, index_bytes, dto, tri_off[m0], tri_off[m1], m0, m1, nt, n_mesh, *params, d, ws, s);
        if (rc != G3D_OK) return bail(rc);
    }
    rc = tdf_solve(dto, nt, n_mesh, *params, ws, s);
    if (rc != G3D_OK) return bail(rc);
    const size_t gv = (size_t)nvox;
    for (int ch = 0; ch < nch; ++ch) {
        const int m0 = chunk_lo(ch), m1 = chunk_lo(ch + 1), nm = m1 - m0;
        rc = tdf_finalize(dvo, dto, nt, n_mesh, m0, m1, *params, d, ws, s);
        if (rc != G3D_OK) return bail(rc);
        G3D_TRY_BAIL(cudaEventRecord(ev_fin[ch], s));
        G3D_TRY_BAIL(cudaStreamWaitEvent(c->s_out, ev_fin[ch], 0));
        const size_t gb = (size_t)nm * gv * 4, go = (size_t)m0 * gv;
        if (out->sdf) G3D_TRY_BAIL(cudaMemcpyAsync(out->sdf + go, d.sdf + go, gb, cudaMemcpyDeviceToHost, c->s_out));
        if (out->tdf) G3D_TRY_BAIL(cudaMemcpyAsync(out->tdf + go, d.tdf + go, gb, cudaMemcpyDeviceToHost, c->s_out));
        if (out->tdflog) G3D_TRY_BAIL(cudaMemcpyAsync(out->tdflog + go, d.tdflog + go, gb, cudaMemcpyDeviceToHost, c->s_out));
        if (out->occ_bits) G3D_TRY_BAIL(cudaMemcpyAsync(out->occ_bits + (size_t)m0 * pw, d.occ_bits + (size_t)m0 * pw, (size_t)nm * pw * 4, cudaMemcpyDeviceToHost, c->s_out));
        if (out->occ_f32) G3D_TRY_BAIL(cudaMemcpyAsync(out->occ_f32 + go, d.occ_f32 + go, gb, cudaMemcpyDeviceToHost, c->s_out));
        if (out->closest_tri) G3D_TRY_BAIL(cudaMemcpyAsync(out->closest_tri + go, d.closest_tri + go, gb, cudaMemcpyDeviceToHost, c->s_out));
    }
    G3D_TRY_BAIL(cudaEventRecord(c->ev_kend[seq & 3], s));
    if (out->status) G3D_TRY_BAIL(cudaMemcpyAsync(out->status, d.status, (size_t)n_mesh * 4, cudaMemcpyDeviceToHost, c->s_out));
    G3D_TRY_BAIL(cudaEventRecord(c->ev_done[slot], c->s_out));
    c->busy[slot] = true;
    *ticket = slot;
    return G3D_OK;
#undef G3D_TRY_BAIL
}

int g3d_tdf_batch_host_submit(g3d_context* c, const float* verts, const int64_t* vert_off, const uint32_t* tris,
                              const int64_t* tri_off, int32_t n_mesh, const g3d_tdf_params* params,
                              const g3d_tdf_outputs* out, int32_t* ticket)
{
    return tdf_host_submit(c, verts, vert_off, tris, 4, tri_off, n_mesh, params, out, ticket);
}

int g3d_tdf_batch_host_submit_u16(g3d_context* c, const float* verts, const int64_t* vert_off, const uint16_t* tris,
                                  const int64_t* tri_off, int32_t n_mesh, const g3d_tdf_params* params,
                                  const g3d_tdf_outputs* out, int32_t* ticket)
{
    if (n_mesh > 0 && vert_off)
        for (int m = 0; m < n_mesh; ++m)
            if (vert_off[m + 1] - vert_off[m] > 65536) {
                set_error("tdf_host: mesh %d has %lld vertices; 16-bit indices address 65,536", m, (long long)(vert_off[m + 1] - vert_off[m]));
                return G3D_ERR_INVALID;
            }
    return tdf_host_submit(c, verts, vert_off, tris, 2, tri_off, n_mesh, params, out, ticket);
}

int g3d_tdf_batch_host_wait(g3d_context* c, int32_t ticket)
{
    if (!c || ticket < 0 || ticket > 1 || !c->busy[ticket]) { set_error("tdf_host_wait: no such batch in flight"); return G3D_ERR_INVALID; }
    G3D_CUDA_TRY(cudaSetDevice(c->device));
    c->busy[ticket] = false;
    G3D_CUDA_TRY(cudaEventSynchronize(c->ev_done[ticket]));
    G3D_CUDA_TRY(cudaGetLastError());
    // (both events lie behind this batch's results on the kernel stream; the ring keeps four submissions apart)
    const unsigned seq = c->seq_of[ticket];
    if (seq > 0 && c->seq - seq < 3) {
        float gap_ms = 0.f;
        if (cudaEventElapsedTime(&gap_ms, c->ev_kend[(seq - 1) & 3], c->ev_kbeg[seq & 3]) == cudaSuccess) c->starved = gap_ms > 0.25f;
        else cudaGetLastError();
    }
    return G3D_OK;
}

int g3d_tdf_batch_host(g3d_context* c, const float* verts, const int64_t* vert_off, const uint32_t* tris,
                       const int64_t* tri_off, int32_t n_mesh, const g3d_tdf_params* params,
                       const g3d_tdf_outputs* out)
{
    int32_t ticket = -1;
    int rc = g3d_tdf_batch_host_submit(c, verts, vert_off, tris, tri_off, n_mesh, params, out, &ticket);
    if (rc != G3D_OK) return rc;
    return g3d_tdf_batch_host_wait(c, ticket);
}

int g3d_raycast_batch_host(g3d_context* c, const uint8_t* depth, const float* pose, int32_t n_view, int32_t H,
                           int32_t W, float focal, float cx, float cy, int32_t clamp_max, int32_t dim,
                           uint32_t* occ_bits, float* occ_f32, int64_t* index_map)
{
    if (!c) { set_error("raycast_host: ctx is NULL"); return G3D_ERR_INVALID; }
    if (n_view < 0 || H < 1 || W < 1 || dim < 1) { set_error("raycast_host: bad sizes"); return G3D_ERR_INVALID; }
    if (n_view == 0) return G3D_OK;
    if (!depth || !pose) { set_error("raycast_host: null input"); return G3D_ERR_INVALID; }
    if (c->busy[0] || c->busy[1]) { set_error("raycast_host: a submitted TDF batch is still in flight on this context; wait for it first"); return G3D_ERR_INVALID; }
    G3D_CUDA_TRY(cudaSetDevice(c->device));
    const int64_t nvox = (int64_t)dim * dim * dim, pw = (nvox + 31) / 32;
    const size_t img = (size_t)n_view * H * W;
    auto layout = [&](char* base, uint8_t*& dd, float*& dp, uint32_t*& db, float*& df, int64_t*& di) {
        Carver k(base);
        dd = k.take<uint8_t>(img);
        dp = k.take<float>((size_t)n_view * 16);
        db = occ_bits ? k.take<uint32_t>((size_t)n_view * pw) : nullptr;
        df = occ_f32 ? k.take<float>((size_t)n_view * nvox) : nullptr;
        di = index_map ? k.take<int64_t>((size_t)n_view * 4 * nvox) : nullptr;
        return k.off;
    };
    uint8_t* dd; float* dp; uint32_t* db; float* df; int64_t* di;
    size_t need = layout(nullptr, dd, dp, db, df, di);
    char* base = arena_reserve(c, need);
    if (!base) return G3D_ERR_CUDA;
    layout(base, dd, dp, db, df, di);
    // The depth images are the bulk of the traffic (H*W bytes per view against dim^3/8 bytes of result), so the call
    // is bound by the host->device copy. The views cross PCIe in chunks on the input stream; a chunk's kernel starts
    // when the chunk has landed and its results leave on the output stream while later chunks are still arriving:
    // total time = the copy of the depth bytes + the tail of the last chunk.
    cudaStream_t s = c->stream;
    auto bail = [&](int code) { cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(s); cudaStreamSynchronize(c->s_out); return code; };
#define G3D_TRY_BAIL(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return bail(g3d::cuda_fail(e__, #expr)); } while (0)
    const int nch = n_view >= 2 * kHostChunks ? kHostChunks * 2 : 1;
    const size_t per = (size_t)H * W;
    cudaEvent_t* ev_in = c->ev_in[0];              // 2 x kHostChunks events per direction: both slots' sets
    cudaEvent_t* ev_fin = c->ev_fin[0];
    G3D_TRY_BAIL(cudaMemcpyAsync(dp, pose, (size_t)n_view * 64, cudaMemcpyHostToDevice, c->s_in));
    for (int ch = 0; ch < nch; ++ch) {
        const int v0 = (int)((int64_t)n_view * ch / nch), v1 = (int)((int64_t)n_view * (ch + 1) / nch);
        G3D_TRY_BAIL(cudaMemcpyAsync(dd + v0 * per, depth + v0 * per, (size_t)(v1 - v0) * per, cudaMemcpyHostToDevice, c->s_in));
        cudaEvent_t ein = ch < kHostChunks ? ev_in[ch] : c->ev_in[1][ch - kHostChunks];
        cudaEvent_t efin = ch < kHostChunks ? ev_fin[ch] : c->ev_fin[1][ch - kHostChunks];
        G3D_TRY_BAIL(cudaEventRecord(ein, c->s_in));
        G3D_TRY_BAIL(cudaStreamWaitEvent(s, ein, 0));
        int rc = g3d_raycast_batch(dd + v0 * per, dp + (size_t)v0 * 16, v1 - v0, H, W, focal, cx, cy, clamp_max, dim,
                                   db ? db + (size_t)v0 * pw : nullptr, df ? df + (size_t)v0 * nvox : nullptr,
                                   di ? di + (size_t)v0 * 4 * nvox : nullptr, s);
        if (rc != G3D_OK) return bail(rc);
        G3D_TRY_BAIL(cudaEventRecord(efin, s));
        G3D_TRY_BAIL(cudaStreamWaitEvent(c->s_out, efin, 0));
        const size_t nv = (size_t)(v1 - v0);
        if (occ_bits) G3D_TRY_BAIL(cudaMemcpyAsync(occ_bits + (size_t)v0 * pw, db + (size_t)v0 * pw, nv * pw * 4, cudaMemcpyDeviceToHost, c->s_out));
        if (occ_f32) G3D_TRY_BAIL(cudaMemcpyAsync(occ_f32 + (size_t)v0 * nvox, df + (size_t)v0 * nvox, nv * nvox * 4, cudaMemcpyDeviceToHost, c->s_out));
        if (index_map) G3D_TRY_BAIL(cudaMemcpyAsync(index_map + (size_t)v0 * 4 * nvox, di + (size_t)v0 * 4 * nvox, nv * 4 * nvox * 8, cudaMemcpyDeviceToHost, c->s_out));
    }
    G3D_TRY_BAIL(cudaStreamSynchronize(c->s_out));
    G3D_TRY_BAIL(cudaStreamSynchronize(s));
#undef G3D_TRY_BAIL
    return G3D_OK;
}

void* g3d_host_alloc(size_t bytes)
{
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { cuda_fail(e, "cudaHostAlloc"); return nullptr; }
    return p;
}

void g3d_host_free(void* p) { if (p) cudaFreeHost(p); }

// ----------------------------------------------------------- dfgen driver ----
int g3d_dfgen_geometry(int32_t dim, int32_t padding, int32_t is_unit_cube, const float* bbox_min,
                       const float* bbox_max, g3d_tdf_params* p, float* g2w)
{
    if (!p || !g2w) { set_error("dfgen_geometry: NULL output"); return G3D_ERR_INVALID; }
    if (dim - 2 * padding <= 0 || dim < 1) { set_error("dfgen_geometry: dim - 2*padding must be positive"); return G3D_ERR_INVALID; }
    if (!is_unit_cube && (!bbox_min || !bbox_max)) { set_error("dfgen_geometry: bbox required"); return G3D_ERR_INVALID; }
    // main.cpp:57: float dx = 1.0/(dim - 2*padding)  (double quotient rounded to float)
    const float dx = (float)(1.0 / (double)(dim - 2 * padding));
    int sizes[3];
    for (int a = 0; a < 3; ++a) {
        float lo = is_unit_cube ? -0.5f : bbox_min[a];             // main.cpp:82-85
        float hi = is_unit_cube ? 0.5f : bbox_max[a];
        const float pad = (float)padding * dx;                     // main.cpp:91-93
        lo -= pad;
        hi += pad;
        const float q = (hi - lo) / dx;                            // main.cpp:94, Vec3ui(...) truncates
        sizes[a] = (q >= 0.f && q < 2147483648.f) ? (int)(unsigned int)q : -1;
        p->origin[a] = lo;
    }
    p->ni = sizes[0]; p->nj = sizes[1]; p->nk = sizes[2];
    p->dx = dx;
    p->out_scale = (float)dim;                                     // main.cpp:119
    p->exact_band = 1;
    const float dfix = (float)(1.0 / (double)dim);                 // main.cpp:107
    const float tr = padding != 0 ? (-0.5f - dfix) : -0.5f;        // main.cpp:110-113
    memset(g2w, 0, 16 * sizeof(float));
    g2w[0] = g2w[5] = g2w[10] = dx;
    g2w[3] = g2w[7] = g2w[11] = tr;
    g2w[15] = 1.f;
    if (sizes[0] != dim || sizes[1] != dim || sizes[2] != dim) {
        set_error("dfgen: grid %dx%dx%d != dim %d (the reference asserts here, main.cpp:102)", sizes[0], sizes[1], sizes[2], dim);
        return G3D_ERR_ASSERT;
    }
    return G3D_OK;
}

int g3d_write_df(const char* path, const int32_t* dims, float res, const float* g2w, const float* sdf)
{
    if (!path || !dims || !g2w || !sdf) { set_error("write_df: NULL argument"); return G3D_ERR_INVALID; }
    FILE* f = fopen(path, "wb");
    if (!f) { set_error("write_df: cannot open %s", path); return G3D_ERR_IO; }
    size_t n = (size_t)dims[0] * dims[1] * dims[2];
    bool ok = fwrite(dims, 4, 3, f) == 3 && fwrite(&res, 4, 1, f) == 1 && fwrite(g2w, 4, 16, f) == 16 &&
              fwrite(sdf, 4, n, f) == n;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { set_error("write_df: short write to %s", path); return G3D_ERR_IO; }
    return G3D_OK;
}

int g3d_read_df(const char* path, int32_t* dims, float* res, float* g2w, float* sdf, int64_t capacity)
{
    if (!path || !dims) { set_error("read_df: NULL argument"); return G3D_ERR_INVALID; }
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("read_df: cannot open %s", path); return G3D_ERR_IO; }
    float r = 0, m[16];
    bool ok = fread(dims, 4, 3, f) == 3 && fread(&r, 4, 1, f) == 1 && fread(m, 4, 16, f) == 16;
    if (ok && res) *res = r;
    if (ok && g2w) memcpy(g2w, m, sizeof m);
    if (ok && sdf) {
        // header values are untrusted: bound each dimension before multiplying (three large dims would wrap the product)
        const bool sane = dims[0] >= 0 && dims[1] >= 0 && dims[2] >= 0 && dims[0] <= 65536 && dims[1] <= 65536 && dims[2] <= 65536;
        int64_t n = sane ? (int64_t)dims[0] * dims[1] * dims[2] : -1;          // <= 2^48
        if (!sane || capacity < 0 || n > capacity) {
            fclose(f);
            set_error("read_df: %s holds %lld values, capacity %lld", path, (long long)n, (long long)capacity);
            return G3D_ERR_INVALID;
        }
        ok = fread(sdf, 4, (size_t)n, f) == (size_t)n;
    }
    fclose(f);
    if (!ok) { set_error("read_df: short read from %s", path); return G3D_ERR_IO; }
    return G3D_OK;
}

}  // extern "C"

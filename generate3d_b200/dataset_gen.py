"""Dataset-scale generation (BASELINE configs[4]; SURVEY.md section 8d "Config 5", 8e): the GT half of the data
set -- what CADVoxelization.py:15-39 loops over, one dfgen + one vox2mesh process per model -- for a contiguous shard
of models per rank, chunk by chunk on one GPU, results resident in HBM for the rank's own trainer replica.

For the synthetic benchmark the inputs are produced ON THE DEVICE from the model index (no PCIe, no files): the same
five-box "table" topology as generate3d_b200.synthetic (T = 60 n^2 triangles), with a counter-based hash instead of
numpy's PCG64 for the per-vertex jitter, the U(0.7, 1) scale and the U(0, 2 pi) turn about y, so a model depends on its
global index only -- not on the chunking or on the rank that owns it. torch is used here as an array language for
the INPUT generator; the hot path is g3d_tdf_batch as everywhere else.
"""
import numpy as np
import torch

from . import dfgen, shard, synthetic

_M32 = 0xFFFFFFFF
_templates = {}


def shard_models(n_models, rank, world_size):
    """Contiguous block of model indices owned by `rank`."""
    return shard.shard_range(n_models, rank, world_size)


def _hash32(x):
    """32-bit integer mix on int64 tensors holding values in [0, 2^32)."""
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & _M32
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & _M32
    return x ^ (x >> 16)


def _template(n, device):
    key = (n, str(device))
    if key not in _templates:
        v64, t = _template_f64(n)                                   # unjittered vertices (float64), topology
        _templates[key] = (torch.from_numpy(v64).to(device), torch.from_numpy(t.astype(np.int32)).to(device))
    return _templates[key]


def _template_f64(n):
    vs, ts, base = [], [], 0
    for (x0, x1, y0, y1, z0, z1) in synthetic.TABLE_BOXES:
        lo = np.array([x0, y0, z0])
        ex, ey, ez = np.array([x1 - x0, 0, 0.0]), np.array([0, y1 - y0, 0.0]), np.array([0, 0, z1 - z0])
        for corner, eu, ev in ((lo, ey, ez), (lo + ex, ez, ey), (lo, ez, ex), (lo + ey, ex, ez), (lo, ex, ey), (lo + ez, ey, ex)):
            v, t = synthetic._face_grid(corner, eu, ev, n)
            vs.append(v)
            ts.append(t + base)
            base += v.shape[0]
    return np.concatenate(vs, 0), np.concatenate(ts, 0)


def generate_tables(first_model, count, n=18, device=None, jitter=1e-3):
    """`count` table meshes, models first_model .. first_model+count-1, in the C-ABI batch layout, on `device`:
    dict(verts f32 [count*nv, 3], vert_off i64 [count+1], tris i32 [count*nt, 3] (mesh-local), tri_off i64 [count+1])."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    tv, tt = _template(n, device)
    nv, nt = tv.shape[0], tt.shape[0]
    m = torch.arange(first_model, first_model + count, dtype=torch.int64, device=device)
    hm = _hash32((m + 1) & _M32)                                              # [count]
    e = torch.arange(nv * 3, dtype=torch.int64, device=device)
    u = _hash32((hm[:, None] ^ _hash32(e + 0x9E3779B9)[None, :]) & _M32).to(torch.float64) * (1.0 / 4294967296.0)
    verts = tv[None, :, :] + (u.reshape(count, nv, 3) * (2.0 * jitter) - jitter)
    scale = 0.7 + 0.3 * (_hash32(hm ^ 0xA511E9B3).to(torch.float64) * (1.0 / 4294967296.0))
    ang = (2.0 * np.pi) * (_hash32(hm ^ 0x63D83595).to(torch.float64) * (1.0 / 4294967296.0))
    c, s = torch.cos(ang)[:, None], torch.sin(ang)[:, None]
    x, y, z = verts[..., 0], verts[..., 1], verts[..., 2]
    verts = torch.stack([c * x + s * z, y, -s * x + c * z], -1) * scale[:, None, None]
    return {"verts": verts.to(torch.float32).reshape(count * nv, 3).contiguous(),
            "vert_off": torch.arange(count + 1, dtype=torch.int64, device=device) * nv,
            "tris": tt.repeat(count, 1).contiguous(),
            "tri_off": torch.arange(count + 1, dtype=torch.int64, device=device) * nt}


def voxelize_shard(first_model, count, chunk=1024, n=18, dim=32, trunc=3.0, mode="faithful", device=None,
                   keep=("tdf", "occ_bits"), timing=None):
    """GT grids of models [first_model, first_model + count): the shard of one rank. Inputs are generated on the
    device chunk by chunk; the outputs named in `keep` (any of dfgen.ALL_OUTPUTS, plus 'status' always) stay
    resident as [count, ...] tensors (tdf + packed occupancy: 135 KB per model, 13.5 GB per 100k models).
    timing: optional dict; receives 'tdf_ms' (CUDA-event time of the g3d_tdf_batch calls alone), 'gen_ms' (input
    generation) and 'chunks'."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    names = tuple(dict.fromkeys(tuple(keep) + ("status",)))
    nvox = dim ** 3
    res = {}
    for name in names:
        shape = (count, (nvox + 31) // 32) if name == "occ_bits" else ((count,) if name == "status" else (count, dim, dim, dim))
        res[name] = torch.empty(shape, dtype=dfgen._DTYPES[name], device=device)
    ev = []
    with torch.cuda.device(device):
        for lo in range(0, count, chunk):
            hi = min(lo + chunk, count)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            if timing is not None:
                e0.record()
            b = generate_tables(first_model + lo, hi - lo, n, device)
            if timing is not None:
                e1.record()
            dfgen.tdf_batch(b["verts"], b["vert_off"], b["tris"], b["tri_off"], dim=dim, trunc=trunc, mode=mode,
                            outputs=names, out={k: v[lo:hi] for k, v in res.items()})
            if timing is not None:
                e2.record()
                ev.append((e0, e1, e2))
        if timing is not None:
            torch.cuda.synchronize(device)
            timing["gen_ms"] = sum(a.elapsed_time(b_) for a, b_, _ in ev)
            timing["tdf_ms"] = sum(b_.elapsed_time(c_) for _, b_, c_ in ev)
            timing["chunks"] = len(ev)
    return res

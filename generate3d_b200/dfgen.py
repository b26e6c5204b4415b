"""Host side of the distance-field path: the Python mirror of src/dfgen
(make_level_set3, makelevelset3.h:12-14; the dfgen driver, main.cpp:45-121) and
of the vox2mesh occupancy text writer (vox2mesh/main.cpp:86-104), on top of the
C ABI in include/g3d.h. Grids stay on the GPU as torch tensors."""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import MODE_EXACT, MODE_FAITHFUL, TdfOutputs, TdfParams

_MODES = {"faithful": MODE_FAITHFUL, "exact": MODE_EXACT, MODE_FAITHFUL: MODE_FAITHFUL, MODE_EXACT: MODE_EXACT}
ALL_OUTPUTS = ("sdf", "tdf", "tdflog", "occ_bits", "occ_f32", "closest_tri", "status")
_DTYPES = {"sdf": torch.float32, "tdf": torch.float32, "tdflog": torch.float32, "occ_bits": torch.int32,
           "occ_f32": torch.float32, "closest_tri": torch.int32, "status": torch.int32}


def _device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("generate3d_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _to_dev(x, dtype, device):
    if isinstance(x, np.ndarray):
        if x.dtype == np.uint32:
            x = x.view(np.int32)
        x = torch.from_numpy(np.ascontiguousarray(x))
    if dtype == torch.int32 and x.dtype == torch.uint32:
        x = x.view(torch.int32)
    return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def dfgen_geometry(dim, padding=0, is_unit_cube=1, bbox_min=None, bbox_max=None):
    """main.cpp:45-114: (TdfParams with ni..dx/out_scale filled, grid2world f32[16] row-major)."""
    p = TdfParams()
    g2w = np.zeros(16, np.float32)
    bmin = np.ascontiguousarray(bbox_min, np.float32) if bbox_min is not None else None
    bmax = np.ascontiguousarray(bbox_max, np.float32) if bbox_max is not None else None
    _lib.check(_lib.lib().g3d_dfgen_geometry(
        dim, padding, is_unit_cube, bmin.ctypes.data if bmin is not None else None,
        bmax.ctypes.data if bmax is not None else None, C.byref(p), g2w.ctypes.data))
    return p, g2w


_ws_cache = {}


def _workspace(nbytes, device, stream):
    """Grow-only scratch per (device, stream), so steady-state calls never allocate and two calls enqueued on
    distinct streams never share scratch (thread-safe for distinct streams). A buffer that is outgrown is
    handed back to torch's caching allocator, which keeps it alive until the work queued on it has drained."""
    key = (device, stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        old = _ws_cache.get(key)
        if old is not None:
            old.record_stream(torch.cuda.ExternalStream(stream, device=device) if stream else torch.cuda.default_stream(device))
        _ws_cache[key] = ws
    return ws


def tdf_batch_params(verts, vert_off, tris, tri_off, params, outputs=("sdf",), device=None, out=None):
    """Run g3d_tdf_batch on device tensors with explicit TdfParams. Returns dict name -> tensor.

    verts f32 [V,3], vert_off i64 [n+1], tris i32/u32 [T,3] (mesh-local indices), tri_off i64 [n+1].
    """
    device = _device(device if device is not None else (verts.device if torch.is_tensor(verts) and verts.is_cuda else None))
    with torch.cuda.device(device):
        L = _lib.lib()
        verts = _to_dev(verts, torch.float32, device).reshape(-1, 3)
        tris = _to_dev(tris, torch.int32, device).reshape(-1, 3)
        n_mesh = int(vert_off.shape[0]) - 1
        total_verts, total_tris = int(verts.shape[0]), int(tris.shape[0])
        vert_off = _to_dev(vert_off, torch.int64, device)
        tri_off = _to_dev(tri_off, torch.int64, device)
        ni, nj, nk = params.ni, params.nj, params.nk
        nvox = ni * nj * nk
        res = {} if out is None else dict(out)
        o = TdfOutputs()
        for name in outputs:
            if name not in ALL_OUTPUTS:
                raise ValueError("unknown output %r" % name)
            if name not in res:
                if name == "occ_bits":
                    shape = (n_mesh, (nvox + 31) // 32)
                elif name == "status":
                    shape = (n_mesh,)
                else:
                    shape = (n_mesh, nk, nj, ni)
                res[name] = torch.empty(shape, dtype=_DTYPES[name], device=device)
            setattr(o, name, res[name].data_ptr())
        nbytes = L.g3d_tdf_workspace_bytes(n_mesh, total_verts, total_tris, C.byref(params))
        stream = torch.cuda.current_stream(device).cuda_stream
        ws = _workspace(nbytes, device, stream)
        _lib.check(L.g3d_tdf_batch(verts.data_ptr(), vert_off.data_ptr(), total_verts, tris.data_ptr(),
                                   tri_off.data_ptr(), total_tris, n_mesh, C.byref(params), C.byref(o),
                                   ws.data_ptr(), ws.numel(), stream))
        return res


def tdf_batch(verts, vert_off, tris, tri_off, dim=32, padding=0, is_unit_cube=1, trunc=3.0,
              occ_thresh=1.0, mode="faithful", outputs=("sdf", "tdf", "tdflog", "occ_bits"), device=None,
              out=None):
    """What `dfgen <obj> dim padding unit out.df` + vox2mesh + load_df + apply_log_transform
    produce for a whole batch of meshes, on the GPU: dict of [n_mesh, dim, dim, dim] tensors
    ([z][y][x], voxel units) / packed occupancy words."""
    if is_unit_cube:
        p, _ = dfgen_geometry(dim, padding, 1)
    else:
        # dfgen derives the grid from the bounding box of ITS mesh (main.cpp:66-76): with several meshes in one call
        # a common box would silently give every model a different field than its own dfgen run
        if int(vert_off.shape[0]) - 1 > 1:
            raise ValueError("is_unit_cube=0 places the grid on each mesh's own bounding box: call tdf_batch once per mesh "
                             "(or pass explicit TdfParams to tdf_batch_params)")
        bmin, bmax = _bbox(verts)
        p, _ = dfgen_geometry(dim, padding, 0, bmin, bmax)
    p.trunc, p.occ_thresh, p.mode = float(trunc), float(occ_thresh), _MODES[mode]
    return tdf_batch_params(verts, vert_off, tris, tri_off, p, outputs, device, out)


def _bbox(verts):
    v = verts.detach().cpu().numpy() if torch.is_tensor(verts) else np.asarray(verts)
    v = v.reshape(-1, 3)
    return v.min(0).astype(np.float32), v.max(0).astype(np.float32)


def make_level_set3(tri, x, origin, dx, nx, ny, nz, exact_band=1, return_closest_tri=False, device=None):
    """make_level_set3 (makelevelset3.h:12-14) for one mesh: phi f32 [nz, ny, nx] on the GPU."""
    p = TdfParams()
    p.ni, p.nj, p.nk = int(nx), int(ny), int(nz)
    p.origin[:] = [float(o) for o in origin]
    p.dx, p.out_scale, p.trunc, p.occ_thresh = float(dx), 1.0, float("inf"), 1.0
    p.mode, p.exact_band = MODE_FAITHFUL, int(exact_band)
    nv = int(np.prod(x.shape)) // 3
    nt = int(np.prod(tri.shape)) // 3
    outs = ("sdf", "closest_tri") if return_closest_tri else ("sdf",)
    r = tdf_batch_params(x, np.array([0, nv], np.int64), tri, np.array([0, nt], np.int64), p, outs, device)
    return (r["sdf"][0], r["closest_tri"][0]) if return_closest_tri else r["sdf"][0]


def load_mesh(path):
    """OBJ / PLY -> (verts f32 [nv,3], tris u32 [nt,3]) as LoaderMesh.h:15-84 does (all vertices
    kept, faces in file order)."""
    L = _lib.lib()
    pv, pt, nv, nt = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
    _lib.check(L.g3d_load_mesh(os.fsencode(path), C.byref(pv), C.byref(nv), C.byref(pt), C.byref(nt)))
    try:
        v = np.ctypeslib.as_array(C.cast(pv, C.POINTER(C.c_float)), (nv.value, 3)).copy() if nv.value else np.zeros((0, 3), np.float32)
        t = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_uint32)), (nt.value, 3)).copy() if nt.value else np.zeros((0, 3), np.uint32)
    finally:
        L.g3d_free(pv)
        L.g3d_free(pt)
    return v, t


def write_df(path, sdf, grid2world):
    """save_vox (LoaderVOX.h:48-63). sdf: [nz,ny,nx] tensor/array."""
    a = sdf.detach().cpu().numpy() if torch.is_tensor(sdf) else np.asarray(sdf)
    a = np.ascontiguousarray(a, np.float32)
    dims = np.array(a.shape[::-1], np.int32)
    g = np.ascontiguousarray(grid2world, np.float32)
    _lib.check(_lib.lib().g3d_write_df(os.fsencode(path), dims.ctypes.data, 1.0, g.ctypes.data, a.ctypes.data))


def read_df(path):
    """load_vox (LoaderVOX.h:20-46) -> (sdf f32 [nz,ny,nx], res, grid2world[16])."""
    L = _lib.lib()
    dims = np.zeros(3, np.int32)
    _lib.check(L.g3d_read_df(os.fsencode(path), dims.ctypes.data, None, None, None, 0))
    n = int(dims[0]) * int(dims[1]) * int(dims[2])
    if min(int(d) for d in dims) < 0 or 80 + 4 * n > os.path.getsize(path):      # untrusted header: never allocate on its say-so
        raise _lib.G3DError(-4, "read_df: %s: header dims %s do not fit the file" % (path, dims.tolist()))
    sdf = np.empty((int(dims[2]), int(dims[1]), int(dims[0])), np.float32)
    res = C.c_float()
    g = np.zeros(16, np.float32)
    _lib.check(L.g3d_read_df(os.fsencode(path), dims.ctypes.data, C.byref(res), g.ctypes.data, sdf.ctypes.data, sdf.size))
    return sdf, res.value, g


def occ_rows_gt(occ_zyx, color=(0, 169, 255)):
    """Text rows 'i j k r g b' in the k, j, i loop order of vox2mesh write_txt
    (vox2mesh/main.cpp:86-104; colour = jet(0))."""
    a = occ_zyx.detach().cpu().numpy() if torch.is_tensor(occ_zyx) else np.asarray(occ_zyx)
    k, j, i = np.nonzero(a)
    return "".join("%d %d %d %d %d %d\n" % (x, y, z, *color) for x, y, z in zip(i, j, k))


def vox2mesh(in_df, out_ply, is_unitless=False, redcenter=False, cmap="jet", trunc=1.0, write_ply=True):
    """The vox2mesh executable as a function (vox2mesh/main.cpp:131-203): .df -> <out>.txt (GT occupancy rows) and,
    unless write_ply is False, the cube-per-voxel ASCII PLY. Host only."""
    _lib.check(_lib.lib().g3d_vox2mesh(os.fsencode(in_df), os.fsencode(out_ply), int(bool(is_unitless)), int(bool(redcenter)),
                                       cmap.encode(), float(trunc), int(bool(write_ply))))


def dfgen(mesh_path, dim, padding, is_unit_cube, out_path, mode="faithful", write_occ_txt=False):
    """The dfgen CLI as a function (main.cpp:17-125): mesh file -> .df file. Returns the sdf tensor."""
    v, t = load_mesh(mesh_path)
    r = tdf_batch(v, np.array([0, len(v)], np.int64), t, np.array([0, len(t)], np.int64), dim, padding,
                  is_unit_cube, mode=mode, outputs=("sdf", "occ_f32"))
    _, g2w = dfgen_geometry(dim, padding, is_unit_cube, *(_bbox(v) if not is_unit_cube else (None, None)))
    write_df(out_path, r["sdf"][0], g2w)
    if write_occ_txt:
        with open(os.path.splitext(out_path)[0] + ".txt", "w") as f:
            f.write(occ_rows_gt(r["occ_f32"][0]))
    return r["sdf"][0]

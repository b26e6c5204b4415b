"""Seeded synthetic inputs of the shapes BASELINE.json names.

There is no ShapeNet data in the image, so benchmarks and parity tests run on a
procedural "table": five axis-aligned boxes inside the unit cube, every face an
n x n quad grid split into two triangles (T = 60 n^2: n=9 -> 4,860, n=18 ->
19,440, n=58 -> 201,840), vertices jittered by U(-1e-3, 1e-3). Depth views are
rendered with the camera model of the reference raycast (raytracing.py:27-33,
64-73; renderer.py:125-156 turns the object about x by 15 degrees per view and
stores uint8(depth * 100), renderer.py:34-40).

Pure numpy on purpose: the same arrays feed the CUDA path, the CPU oracle and
the compiled reference, on any machine.
"""
import numpy as np

# (xmin, xmax, ymin, ymax, zmin, zmax): table top + four legs.
TABLE_BOXES = (
    (-0.45, 0.45, 0.15, 0.20, -0.30, 0.30),
    (-0.42, -0.37, -0.40, 0.15, -0.27, -0.22),
    (0.37, 0.42, -0.40, 0.15, -0.27, -0.22),
    (-0.42, -0.37, -0.40, 0.15, 0.22, 0.27),
    (0.37, 0.42, -0.40, 0.15, 0.22, 0.27),
)


def _face_grid(corner, eu, ev, n):
    """(n+1)^2 vertices and 2 n^2 triangles of one quad face."""
    s = np.linspace(0.0, 1.0, n + 1)
    uu, vv = np.meshgrid(s, s, indexing="ij")
    verts = corner[None, :] + uu.reshape(-1, 1) * eu[None, :] + vv.reshape(-1, 1) * ev[None, :]
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a = (i * (n + 1) + j).ravel()
    b = a + (n + 1)
    tris = np.concatenate([np.stack([a, b, b + 1], 1), np.stack([a, b + 1, a + 1], 1)], 0)
    return verts, tris


def table_mesh(n=9, seed=0, scale=1.0, rot_y=0.0, jitter=1e-3):
    """Return (verts f32 [30 (n+1)^2, 3], tris u32 [60 n^2, 3])."""
    vs, ts, base = [], [], 0
    for (x0, x1, y0, y1, z0, z1) in TABLE_BOXES:
        lo = np.array([x0, y0, z0])
        ex, ey, ez = np.array([x1 - x0, 0, 0.0]), np.array([0, y1 - y0, 0.0]), np.array([0, 0, z1 - z0])
        faces = (
            (lo, ey, ez), (lo + ex, ez, ey),      # -x, +x
            (lo, ez, ex), (lo + ey, ex, ez),      # -y, +y
            (lo, ex, ey), (lo + ez, ey, ex),      # -z, +z
        )
        for corner, eu, ev in faces:
            v, t = _face_grid(corner, eu, ev, n)
            vs.append(v)
            ts.append(t + base)
            base += v.shape[0]
    verts = np.concatenate(vs, 0)
    tris = np.concatenate(ts, 0)
    rng = np.random.default_rng(seed)
    verts = verts + rng.uniform(-jitter, jitter, size=verts.shape)
    if rot_y != 0.0:
        c, s = np.cos(rot_y), np.sin(rot_y)
        rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        verts = verts @ rot.T
    verts = verts * scale
    return np.ascontiguousarray(verts, dtype=np.float32), np.ascontiguousarray(tris, dtype=np.uint32)


def table_batch(n_mesh, n=18, first_seed=0):
    """Ragged batch in the C-ABI layout: verts [sum nv,3], vert_off [n+1], tris [sum T,3]
    (indices local to each mesh), tri_off [n+1]. Model m uses seed first_seed+m, a
    U(0.7,1) scale and a U(0,2pi) turn about y so that shards are not identical."""
    vs, ts, voff, toff = [], [], [0], [0]
    for m in range(n_mesh):
        seed = first_seed + m
        prm = np.random.default_rng(1_000_003 * (seed + 1))
        v, t = table_mesh(n, seed, scale=prm.uniform(0.7, 1.0), rot_y=prm.uniform(0, 2 * np.pi))
        vs.append(v)
        ts.append(t)
        voff.append(voff[-1] + v.shape[0])
        toff.append(toff[-1] + t.shape[0])
    return (np.concatenate(vs, 0), np.asarray(voff, np.int64), np.concatenate(ts, 0),
            np.asarray(toff, np.int64))


def view_pose(k, cam_depth=1.4, cam_height=0.1, deg_per_view=15.0):
    """4x4 f32 world->camera matrix of view k *after* load_pose's patch
    (dataset_loader.py:70-81: [2,3] = -cam_depth, [1,3] = 0.1)."""
    a = np.deg2rad(deg_per_view * k)
    c, s = np.cos(a), np.sin(a)
    pose = np.eye(4)
    pose[:3, :3] = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    pose = pose.astype(np.float32)
    pose[2, 3] = -cam_depth
    pose[1, 3] = cam_height
    return pose


def pose_json_payload(k, deg_per_view=15.0):
    """The 16 numbers renderer.py:43-60 writes (column-major, before the patch)."""
    a = np.deg2rad(deg_per_view * k)
    c, s = np.cos(a), np.sin(a)
    pose = np.eye(4)
    pose[:3, :3] = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    return [float(x) for x in pose.T.reshape(-1)]


def _cross2(a, b):
    return a[0] * b[1] - a[1] * b[0]


def _hull_mask(pts, H, W):
    """Boolean H x W mask of pixel centres inside the convex hull of 2-D pts."""
    pts = np.unique(np.round(pts, 6), axis=0)
    if pts.shape[0] < 3:
        return np.zeros((H, W), bool)
    # Andrew monotone chain
    pts = pts[np.lexsort((pts[:, 1], pts[:, 0]))]

    def half(seq):
        out = []
        for p in seq:
            while len(out) >= 2 and _cross2(out[-1] - out[-2], p - out[-2]) <= 0:
                out.pop()
            out.append(p)
        return out

    hull = half(pts)[:-1] + half(pts[::-1])[:-1]
    if len(hull) < 3:
        return np.zeros((H, W), bool)
    vv, uu = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    inside = np.ones((H, W), bool)
    for a, b in zip(hull, hull[1:] + hull[:1]):
        inside &= ((b[0] - a[0]) * (vv - a[1]) - (b[1] - a[1]) * (uu - a[0])) >= 0
    return inside


def depth_view(k, H=256, W=256, focal=None, boxes=TABLE_BOXES, depth_factor=100):
    """uint8 H x W depth image of the table seen from view k (0 = background)."""
    focal = 525.0 * W / 512.0 if focal is None else focal
    cx, cy = W / 2.0, H / 2.0
    pose = view_pose(k).astype(np.float64)
    depth = np.zeros((H, W), np.float64)
    for (x0, x1, y0, y1, z0, z1) in boxes:
        corners = np.array([[x, y, z, 1.0] for x in (x0, x1) for y in (y0, y1) for z in (z0, z1)])
        cam = corners @ pose.T
        z = np.abs(cam[:, 2])
        uv = np.stack([cam[:, 0] * focal / z + cx, -cam[:, 1] * focal / z + cy], 1)
        mask = _hull_mask(uv, H, W)
        zc = float(z.mean())
        upd = mask & ((depth == 0) | (zc < depth))
        depth[upd] = zc
    return np.clip(depth * depth_factor, 0, 255).astype(np.uint8)


def depth_views(n_view=20, H=256, W=256):
    """(depth u8 [n_view,H,W], poses f32 [n_view,4,4]) for views 0..n_view-1."""
    d = np.stack([depth_view(k, H, W) for k in range(n_view)], 0)
    p = np.stack([view_pose(k) for k in range(n_view)], 0)
    return d, p

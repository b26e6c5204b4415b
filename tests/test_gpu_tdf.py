"""GPU parity tests of the distance-field path: CUDA (through the C ABI) vs the CPU oracle,
the compiled reference and the golden fixtures. Bit-exact for faithful mode."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from generate3d_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _oracle_full(oracle, v, t, dim):
    g = oracle.dfgen_geometry(dim)
    r = oracle.make_level_set3(t, v, g["origin"], g["dx"], dim, dim, dim)
    r["sdf"] = r["phi"] * np.float32(dim)
    return r


def _run(dfgen, meshes, dim=32, mode="faithful", outputs=("sdf", "tdf", "tdflog", "occ_bits", "occ_f32", "closest_tri", "status")):
    voff, toff = [0], [0]
    for v, t in meshes:
        voff.append(voff[-1] + len(v))
        toff.append(toff[-1] + len(t))
    V = np.concatenate([m[0].reshape(-1, 3) for m in meshes], 0).astype(np.float32)
    T = np.concatenate([m[1].reshape(-1, 3) for m in meshes], 0).astype(np.uint32)
    r = dfgen.tdf_batch(V, np.asarray(voff, np.int64), T, np.asarray(toff, np.int64), dim=dim, mode=mode, outputs=outputs)
    torch.cuda.synchronize()
    return {k: x.cpu().numpy() for k, x in r.items()}


@pytest.mark.parametrize("n,dim,seed", [(9, 32, 0), (9, 32, 1), (18, 32, 2), (5, 64, 3), (3, 16, 4)])
def test_faithful_bit_exact_vs_oracle(g3d, oracle, n, dim, seed):
    from generate3d_b200 import dfgen
    v, t = S.table_mesh(n, seed)
    want = _oracle_full(oracle, v, t, dim)
    got = _run(dfgen, [(v, t)], dim)
    assert got["status"][0] == 0
    assert np.array_equal(_bits(got["sdf"][0]), _bits(want["sdf"]))
    assert np.array_equal(got["closest_tri"][0], want["closest_tri"])
    occ = oracle.occupancy(want["sdf"])
    assert np.array_equal(got["occ_f32"][0].astype(np.uint8), occ)
    bits = np.unpackbits(got["occ_bits"][0].view(np.uint8), bitorder="little")[:dim ** 3].reshape(dim, dim, dim)
    assert np.array_equal(bits, occ)
    tdf = oracle.load_df_clamp(want["sdf"], 3.0)
    assert np.array_equal(_bits(got["tdf"][0]), _bits(tdf))
    assert np.abs(got["tdflog"][0] - oracle.apply_log_transform(tdf)).max() <= 1e-5      # north-star tolerance
    if oracle.ref_available():
        g = oracle.dfgen_geometry(dim)
        ref = oracle.ref_make_level_set3(t, v, g["origin"], g["dx"], dim, dim, dim) * np.float32(dim)
        assert np.array_equal(_bits(got["sdf"][0]), _bits(ref))


def test_faithful_ragged_batch_with_bad_and_empty_meshes(g3d, oracle):
    from generate3d_b200 import dfgen
    rng = np.random.default_rng(3)
    meshes = [S.table_mesh(4, 10), S.table_mesh(7, 11, scale=0.8, rot_y=0.7), S.table_mesh(2, 12)]
    # degenerate triangles: repeated vertices, collinear, zero area, far outside the cube
    v, t = S.table_mesh(3, 13)
    t = t.copy()
    t[5] = [t[5][0], t[5][0], t[5][1]]
    t[9] = [t[9][2], t[9][2], t[9][2]]
    v = np.concatenate([v, np.float32([[3, 3, 3], [3.1, 3, 3], [3, 3.2, 3], [0, 0, 0], [0.1, 0.1, 0.1], [0.2, 0.2, 0.2]])], 0)
    nv = len(v)
    t = np.concatenate([t, np.uint32([[nv - 6, nv - 5, nv - 4], [nv - 3, nv - 2, nv - 1]])], 0)
    meshes.append((v, t))
    empty = (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    meshes.append(empty)
    bv, bt = S.table_mesh(2, 14)
    bt = bt.copy()
    bt[3, 1] = len(bv) + 7                                  # out-of-range index
    meshes.append((bv, bt))
    meshes.append(S.table_mesh(5, 15))                      # a good mesh after the bad ones
    got = _run(dfgen, meshes, 32)
    assert got["status"].tolist() == [0, 0, 0, 0, 2, 1, 0]
    for m in (0, 1, 2, 3, 6):
        want = _oracle_full(oracle, meshes[m][0], meshes[m][1], 32)
        assert np.array_equal(_bits(got["sdf"][m]), _bits(want["sdf"])), m
        assert np.array_equal(got["closest_tri"][m], want["closest_tri"]), m
    # empty mesh: the reference's initial field (ni+nj+nk)*dx scaled, no triangle anywhere
    assert np.all(got["sdf"][4] == np.float32(96 * 0.03125 * 32)) and np.all(got["closest_tri"][4] == -1)
    # mesh with one bad face = the same mesh without that face
    want = _oracle_full(oracle, bv, np.delete(bt, 3, 0), 32)
    assert np.array_equal(_bits(got["sdf"][5]), _bits(want["sdf"]))


def test_faithful_random_soup_and_noncubic_grid(g3d, oracle):
    from generate3d_b200 import dfgen, _lib
    rng = np.random.default_rng(5)
    v = rng.uniform(-0.6, 0.6, (300, 3)).astype(np.float32)     # pokes outside the cube
    t = rng.integers(0, 300, (500, 3)).astype(np.uint32)
    origin = np.float32([-0.55, -0.5, -0.45])
    ni, nj, nk = 37, 25, 19                                     # not multiples of 32: atomicOr occupancy path
    want = oracle.make_level_set3(t, v, origin, 0.04, ni, nj, nk)
    phi, ct = dfgen.make_level_set3(t, v, origin, 0.04, ni, nj, nk, return_closest_tri=True)
    assert np.array_equal(_bits(phi.cpu().numpy()), _bits(want["phi"]))
    assert np.array_equal(ct.cpu().numpy(), want["closest_tri"])
    p = _lib.TdfParams()
    p.ni, p.nj, p.nk = ni, nj, nk
    p.origin[:] = [float(x) for x in origin]
    p.dx, p.out_scale, p.trunc, p.occ_thresh, p.mode, p.exact_band = 0.04, 25.0, 3.0, 1.0, 0, 1
    r = dfgen.tdf_batch_params(v, np.array([0, 300]), t, np.array([0, 500]), p, ("sdf", "occ_bits", "occ_f32"))
    occ = oracle.occupancy(want["phi"] * np.float32(25.0))
    assert np.array_equal(r["occ_f32"][0].cpu().numpy().astype(np.uint8), occ)
    bits = np.unpackbits(r["occ_bits"][0].cpu().numpy().view(np.uint8), bitorder="little")[:ni * nj * nk]
    assert np.array_equal(bits.reshape(nk, nj, ni), occ)


def test_stages_init_and_parity(g3d, oracle):
    """Unit-level parity: band init (phi, closest_tri) and intersection parity, via exact_band/sign."""
    from generate3d_b200 import dfgen
    v, t = S.table_mesh(6, 21)
    g = oracle.dfgen_geometry(32)
    full = oracle.make_level_set3(t, v, g["origin"], g["dx"], 32, 32, 32)
    phi = dfgen.make_level_set3(t, v, g["origin"], g["dx"], 32, 32, 32).cpu().numpy()
    # sign pattern == running parity of the oracle's intersection counts (makelevelset3.cpp:174-182)
    par = (np.cumsum(full["isect"], axis=2) % 2) == 1
    assert np.array_equal(np.signbit(phi), par)
    # exact_band = 2 changes the init boxes; still bit-equal
    want2 = oracle.make_level_set3(t, v, g["origin"], g["dx"], 32, 32, 32, exact_band=2)
    phi2 = dfgen.make_level_set3(t, v, g["origin"], g["dx"], 32, 32, 32, exact_band=2).cpu().numpy()
    assert np.array_equal(_bits(phi2), _bits(want2["phi"]))


def test_golden_df_bytes_and_loader_dropins(g3d, golden_dir, tmp_path):
    from generate3d_b200 import dataset_loader, dfgen, losses
    z = np.load(os.path.join(golden_dir, "loader.npz"))
    obj = tmp_path / "model_normalized.obj"
    with open(obj, "w") as f:
        for p in z["verts"]:
            f.write("v %.9g %.9g %.9g\n" % tuple(p))
        for tri in z["tris"]:
            f.write("f %d %d %d\n" % tuple(tri + 1))
    out = tmp_path / "m__0__.df"
    dfgen.dfgen(str(obj), 32, 0, 1, str(out), write_occ_txt=True)
    assert out.read_bytes() == bytes(z["df_file"])                       # byte-identical to the reference pipeline
    df3 = dataset_loader.load_df(str(out), 3.0)
    assert df3.is_cuda and tuple(df3.shape) == (1, 32, 32, 32)
    assert np.array_equal(df3.cpu().numpy(), z["tdf3"])
    occ = dataset_loader.load_occ(str(tmp_path / "m__0__.txt"))
    assert np.array_equal(occ.cpu().numpy().astype(np.uint8), z["occ_gt"])
    tl = losses.apply_log_transform(df3)
    assert np.abs(tl.cpu().numpy() - z["tdflog3"]).max() <= 1e-5
    # the C++ CLI writes the same bytes
    cli = os.path.join(os.path.dirname(g3d.LIB_PATH), "bin", "dfgen")
    out2 = tmp_path / "cli.df"
    subprocess.check_call([cli, str(obj), "32", "0", "1", str(out2)], stdout=subprocess.DEVNULL)
    assert out2.read_bytes() == bytes(z["df_file"])
    assert subprocess.call([cli, str(obj), "32"], stdout=subprocess.DEVNULL) == 255      # usage, exit(-1)


def test_host_buffer_abi_matches_device_abi(g3d, oracle):
    v, t = S.table_mesh(5, 31)
    L = g3d.lib()
    ctx = C.c_void_p()
    g3d.check(L.g3d_context_create(0, C.byref(ctx)))
    p = g3d.TdfParams()
    g2w = np.zeros(16, np.float32)
    g3d.check(L.g3d_dfgen_geometry(32, 0, 1, None, None, C.byref(p), g2w.ctypes.data))
    p.trunc, p.occ_thresh, p.mode = 3.0, 1.0, 0
    sdf = np.empty((2, 32, 32, 32), np.float32)
    bits = np.empty((2, 1024), np.uint32)
    status = np.empty(2, np.int32)
    o = g3d.TdfOutputs()
    o.sdf, o.occ_bits, o.status = sdf.ctypes.data, bits.ctypes.data, status.ctypes.data
    V = np.concatenate([v, v * np.float32(0.5)], 0)
    T = np.concatenate([t, t], 0)
    voff = np.array([0, len(v), 2 * len(v)], np.int64)
    toff = np.array([0, len(t), 2 * len(t)], np.int64)
    g3d.check(L.g3d_tdf_batch_host(ctx, V.ctypes.data, voff.ctypes.data, T.ctypes.data, toff.ctypes.data, 2,
                                   C.byref(p), C.byref(o)))
    L.g3d_context_destroy(ctx)
    assert status.tolist() == [0, 0]
    for m, vv in enumerate((v, v * np.float32(0.5))):
        want = _oracle_full(oracle, vv, t, 32)
        assert np.array_equal(_bits(sdf[m]), _bits(want["sdf"]))
        occ = np.unpackbits(bits[m].view(np.uint8), bitorder="little").reshape(32, 32, 32)
        assert np.array_equal(occ, oracle.occupancy(want["sdf"]))


def test_full_size_batch_properties(g3d, oracle):
    """BASELINE config 3 shape (n=18, 19,440 triangles) at a reduced count: every mesh of the batch equals the
    single-mesh result (batch independence), spot-checked against the oracle."""
    from generate3d_b200 import dfgen
    V, voff, T, toff = S.table_batch(24, 18, first_seed=100)
    r = dfgen.tdf_batch(V, voff, T, toff, outputs=("sdf", "status"))
    sdf = r["sdf"].cpu().numpy()
    assert not r["status"].any()
    for m in (0, 7, 23):
        v, t = V[voff[m]:voff[m + 1]], T[toff[m]:toff[m + 1]]
        one = dfgen.tdf_batch(v, np.array([0, len(v)]), t, np.array([0, len(t)]), outputs=("sdf",))["sdf"][0].cpu().numpy()
        assert np.array_equal(_bits(one), _bits(sdf[m]))
    want = _oracle_full(oracle, V[voff[7]:voff[8]], T[toff[7]:toff[8]], 32)
    assert np.array_equal(_bits(sdf[7]), _bits(want["sdf"]))


@pytest.mark.parametrize("n,dim,seed", [(5, 32, 40), (3, 16, 41), (4, 20, 42)])
def test_exact_mode_vs_brute_force_oracle(g3d, oracle, n, dim, seed):
    """Exact mode = true minimum of the reference's distance function over all triangles wherever it can matter:
    outside voxels within the truncation, and every inside voxel; 'far' elsewhere. Bit-exact, ties to the lowest
    triangle index. Occupancy (|d| <= 1) equals the reference's (its band is exact there)."""
    from generate3d_b200 import dfgen
    v, t = S.table_mesh(n, seed)
    g = oracle.dfgen_geometry(dim)
    bf, bt = oracle.brute_force(t, v, g["origin"], g["dx"], dim, dim, dim)
    faithful = _oracle_full(oracle, v, t, dim)
    got = _run(dfgen, [(v, t)], dim, mode="exact")
    assert got["status"][0] == 0
    sdf = got["sdf"][0]
    inside = np.signbit(faithful["sdf"])
    assert np.array_equal(np.signbit(sdf), inside)                           # same x-ray parity signs
    want = bf * np.float32(dim)
    far = np.float32(3 * dim * g["dx"] * dim)
    near = inside | (want <= 3.0)
    assert np.array_equal(_bits(np.abs(sdf)[near]), _bits(want[near]))       # exact where it matters
    assert np.array_equal(got["closest_tri"][0][near], bt[near])
    beyond = ~inside & (want > 3.0 * 1.001)
    assert np.all(sdf[beyond] == far) and np.all(got["closest_tri"][0][beyond] == -1)
    assert np.array_equal(got["occ_f32"][0].astype(np.uint8), oracle.occupancy(faithful["sdf"]))
    assert np.all(np.abs(sdf)[near] <= np.abs(faithful["sdf"])[near])        # exact <= reference-faithful
    tdf = oracle.load_df_clamp(np.where(near, np.copysign(want, faithful["sdf"]), far), 3.0)
    assert np.array_equal(_bits(got["tdf"][0]), _bits(tdf))


def test_exact_mode_untruncated_and_ragged(g3d, oracle):
    """trunc = inf turns exact mode into a full brute force (branch-and-bound only); ragged batch + bad mesh."""
    from generate3d_b200 import dfgen, _lib
    meshes = [S.table_mesh(3, 50), S.table_mesh(2, 51, scale=0.7, rot_y=1.0)]
    bv, bt_ = S.table_mesh(2, 52)
    bt_ = bt_.copy(); bt_[1, 2] = 99999
    meshes.append((bv, bt_))
    p, _ = dfgen.dfgen_geometry(16)
    p.trunc, p.occ_thresh, p.mode = float("inf"), 1.0, _lib.MODE_EXACT
    voff, toff = [0], [0]
    for v, t in meshes:
        voff.append(voff[-1] + len(v)); toff.append(toff[-1] + len(t))
    V = np.concatenate([m[0] for m in meshes]); T = np.concatenate([m[1] for m in meshes])
    r = dfgen.tdf_batch_params(V, np.asarray(voff), T, np.asarray(toff), p, ("sdf", "closest_tri", "status"))
    assert r["status"].cpu().tolist() == [0, 0, 1]
    g = oracle.dfgen_geometry(16)
    for m, (v, t) in enumerate(meshes):
        if m == 2:
            t = np.delete(t, 1, 0)
        bf, bt = oracle.brute_force(t, v, g["origin"], g["dx"], 16, 16, 16)
        assert np.array_equal(_bits(np.abs(r["sdf"][m].cpu().numpy())), _bits(bf * np.float32(16))), m
        if m != 2:
            assert np.array_equal(r["closest_tri"][m].cpu().numpy(), bt), m


def test_cadvoxelization_dropin_writes_reference_files(g3d, oracle, tmp_path):
    """The batched CADVoxelization mirror leaves the same two files per model the reference pipeline does."""
    from generate3d_b200 import CADVoxelization as CV
    root = tmp_path / "shapenet"
    want = {}
    for k in range(5):
        v, t = S.table_mesh(3, 60 + k, scale=0.9)
        d = root / "04379243" / ("model%02d" % k) / "models"
        d.mkdir(parents=True)
        with open(d / "model_normalized.obj", "w") as f:
            for p in v:
                f.write("v %.9g %.9g %.9g\n" % tuple(p))
            for tri in t:
                f.write("f %d %d %d\n" % tuple(tri + 1))
        want["model%02d" % k] = oracle.dfgen(t, v, 32)
    bad = root / "04379243" / "broken" / "models"
    bad.mkdir(parents=True)
    (bad / "model_normalized.obj").write_text("# nothing here\n")
    params = {"shapenet": str(root), "shapenet_voxelized": str(tmp_path / "vox") + "/"}
    models = CV.list_models(params)
    assert len(models) == 6
    keep = {}
    done, failed = CV.voxelize_models(models, params["shapenet_voxelized"], chunk=4, keep=keep)
    assert done == 5 and [f[0] for f in failed] == ["broken"]
    for mid, g in want.items():
        ref = tmp_path / "ref.df"
        oracle.write_df(str(ref), g["sdf"], g["grid2world"])
        out = tmp_path / "vox" / "04379243" / (mid + "__0__.df")
        assert out.read_bytes() == ref.read_bytes()
        assert (tmp_path / "vox" / "04379243" / (mid + "__0__.txt")).read_text() == oracle.gt_occ_txt_rows(g["sdf"])
        assert keep[mid]["tdf"].is_cuda
    done2, _ = CV.voxelize_models(models, params["shapenet_voxelized"], chunk=4)      # resume: nothing left but the broken one
    assert done2 == 0


def test_cabi_error_paths(g3d):
    """Bad arguments come back as error codes with a message, never as a crash or a silent fallback."""
    L = g3d.lib()
    p = g3d.TdfParams()
    g2w = np.zeros(16, np.float32)
    g3d.check(L.g3d_dfgen_geometry(32, 0, 1, None, None, C.byref(p), g2w.ctypes.data))
    p.trunc, p.occ_thresh, p.mode = 3.0, 1.0, 0
    v = torch.zeros((3, 3), device="cuda")
    t = torch.zeros((1, 3), dtype=torch.int32, device="cuda")
    off = torch.tensor([0, 3], device="cuda"); toff = torch.tensor([0, 1], device="cuda")
    sdf = torch.empty((1, 32, 32, 32), device="cuda")
    o = g3d.TdfOutputs(); o.sdf = sdf.data_ptr()
    need = L.g3d_tdf_workspace_bytes(1, 3, 1, C.byref(p))
    assert need > 0
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    args = lambda wsb, mode=0: (v.data_ptr(), off.data_ptr(), 3, t.data_ptr(), toff.data_ptr(), 1, 1, C.byref(p), C.byref(o), ws.data_ptr(), wsb, None)
    assert L.g3d_tdf_batch(*args(need)) == 0
    assert L.g3d_tdf_batch(*args(need // 2)) == -3 and b"workspace" in L.g3d_last_error()      # G3D_ERR_WORKSPACE
    p.mode = 7
    assert L.g3d_tdf_batch(*args(need)) == -1 and b"mode" in L.g3d_last_error()
    p.mode, p.ni = 0, 0
    assert L.g3d_tdf_batch(*args(need)) == -1
    p.ni = 32
    # the evaluation queues carry global triangle numbers in 32 bits: 2^32 - 1 triangles in one call are one too many
    # (the per-MESH limit of 2^22 - 2, 2^27 - 2 in exact mode, is enforced on the device: G3D_MESH_TOO_MANY_TRIS)
    assert L.g3d_tdf_batch(v.data_ptr(), off.data_ptr(), 3, t.data_ptr(), toff.data_ptr(), (1 << 32) - 1, 1, C.byref(p),
                           C.byref(o), ws.data_ptr(), need, None) == -1 and b"triangles" in L.g3d_last_error()
    assert L.g3d_tdf_batch(v.data_ptr(), off.data_ptr(), 3, t.data_ptr(), toff.data_ptr(), (1 << 27) + 5, 1, C.byref(p),
                           C.byref(o), ws.data_ptr(), need, None) == -3                       # accepted, but this workspace is too small
    assert L.g3d_tdf_batch(None, off.data_ptr(), 3, t.data_ptr(), toff.data_ptr(), 1, 1, C.byref(p), C.byref(o), ws.data_ptr(), need, None) == -1
    assert L.g3d_raycast_batch(None, None, 1, 64, 64, 525.0, 32.0, 32.0, 511, 32, None, None, None, None) == -1
    assert L.g3d_raycast_batch(v.data_ptr(), v.data_ptr(), 1, 4096, 4096, 525.0, 32.0, 32.0, 511, 32, sdf.data_ptr(), None, None, None) == -1
    assert b"shared memory" in L.g3d_last_error()
    assert L.g3d_occ_to_df_batch(sdf.data_ptr(), 1, 100, 0, 3.0, 3.0, sdf.data_ptr(), None) == -1
    torch.cuda.synchronize()                                     # nothing above poisoned the context
    assert L.g3d_tdf_batch(*args(need)) == 0
    torch.cuda.synchronize()


def test_faithful_single_huge_triangle_and_outside_mesh(g3d, oracle):
    """One triangle spanning the whole grid (band box = every node, the integer-division path of the init) and a
    mesh lying completely outside the unit cube (boxes clamp to the boundary layer)."""
    from generate3d_b200 import dfgen
    g = oracle.dfgen_geometry(32)
    big_v = np.float32([[-2, -2, 0.01], [2, -2, 0.02], [0, 3, -0.03], [0.1, 0.1, 0.4], [0.2, 0.1, 0.4], [0.1, 0.3, 0.45]])
    big_t = np.uint32([[0, 1, 2], [3, 4, 5]])
    out_v, out_t = S.table_mesh(2, 70)
    out_v = out_v + np.float32([1.7, 0, 0])
    for v, t in ((big_v, big_t), (out_v, out_t)):
        want = oracle.make_level_set3(t, v, g["origin"], g["dx"], 32, 32, 32)
        phi, ct = dfgen.make_level_set3(t, v, g["origin"], g["dx"], 32, 32, 32, return_closest_tri=True)
        assert np.array_equal(_bits(phi.cpu().numpy()), _bits(want["phi"]))
        assert np.array_equal(ct.cpu().numpy(), want["closest_tri"])


@pytest.mark.parametrize("count,n", [(70, 3), (640, 1)])
def test_host_abi_chunked_path_equals_device_abi(g3d, count, n):
    """>= 64 meshes: the host front end stages triangles in 4 chunks on a copy stream and returns results per chunk.
    Results must equal the one-shot device API bit for bit (faithful and exact), also when the same context is
    used again (arena and event reuse); 640 meshes is enough for the tiled band init to be chosen per chunk."""
    from generate3d_b200 import dfgen
    V, voff, T, toff = S.table_batch(count, n, first_seed=300)
    L = g3d.lib()
    ctx = C.c_void_p()
    g3d.check(L.g3d_context_create(0, C.byref(ctx)))
    for mode in (0, 1, 0):
        p = g3d.TdfParams()
        g2w = np.zeros(16, np.float32)
        g3d.check(L.g3d_dfgen_geometry(32, 0, 1, None, None, C.byref(p), g2w.ctypes.data))
        p.trunc, p.occ_thresh, p.mode = 3.0, 1.0, mode
        sdf = np.empty((count, 32, 32, 32), np.float32)
        ct = np.empty((count, 32, 32, 32), np.int32)
        status = np.empty(count, np.int32)
        o = g3d.TdfOutputs()
        o.sdf, o.closest_tri, o.status = sdf.ctypes.data, ct.ctypes.data, status.ctypes.data
        g3d.check(L.g3d_tdf_batch_host(ctx, V.ctypes.data, voff.ctypes.data, T.ctypes.data, toff.ctypes.data, count,
                                       C.byref(p), C.byref(o)))
        r = dfgen.tdf_batch(V, voff, T, toff, mode="faithful" if mode == 0 else "exact", outputs=("sdf", "closest_tri", "status"))
        assert not status.any()
        assert np.array_equal(_bits(sdf), _bits(r["sdf"].cpu().numpy()))
        assert np.array_equal(ct, r["closest_tri"].cpu().numpy())
    L.g3d_context_destroy(ctx)


@pytest.mark.parametrize("tiled,sweep_threads,bitmap", [(None, None, None), ("1", None, None), ("1", "128", None),
                                                        ("0", "256", None), ("1", "128", "0"), (None, None, "0"),
                                                        ("1", "128", "spec"), (None, None, "spec"), ("1", "256", "split3")])
def test_randomised_parity_campaign(g3d, tiled, sweep_threads, bitmap):
    """A short run of tools/fuzz_parity.py (random soups, degenerate / huge triangles, random grids and poses), once
    with the kernel choice left to the library (single meshes: scatter band init, 1,024-thread sweeps) and then with
    the batch-sized kernels forced onto every case: tiled band init, 128- and 256-thread sweep CTAs (two rounds of a
    level per evaluation queue); with the small-grid sweep kernel (activity bits + speculation) switched off (the
    plain kernel that larger grids take), made to run all 16 sweeps and to speculate in every one of them ("spec"), and
    taking over after sweep 2 ("split3")."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.pop("G3D_INIT_TILED", None)
    env.pop("G3D_SWEEP_THREADS", None)
    env.pop("G3D_SWEEP_BITMAP", None)
    for k in ("G3D_SWEEP_SPEC", "G3D_SWEEP_SPLIT"):
        env.pop(k, None)
    if bitmap == "0":
        env["G3D_SWEEP_BITMAP"] = "0"
    elif bitmap == "spec":
        env["G3D_SWEEP_SPEC"], env["G3D_SWEEP_SPLIT"] = "-1", "0"
    elif bitmap == "split3":
        env["G3D_SWEEP_SPEC"], env["G3D_SWEEP_SPLIT"] = "-1", "3"
    if tiled:
        env["G3D_INIT_TILED"] = tiled
    if sweep_threads:
        env["G3D_SWEEP_THREADS"] = sweep_threads
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "--meshes", "32",
                          "--poses", "100" if tiled is None else "0", "--seed", "5"],
                         capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and "fuzz ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("tiled", ["0", "1", None])
def test_band_init_kernels_match_oracle_on_ragged_multiblock_batch(g3d, oracle, tiled, monkeypatch):
    """Both band-init kernels (scatter with global atomics / tiled with the keys of a 16^3 block in shared memory) on
    a batch whose grid is not a multiple of the tile and whose triangles range from sub-voxel to larger than the
    grid; phi and closest_tri bit for bit against the oracle for every mesh. `None` leaves the choice to the library
    (48 meshes x 12 blocks: tiled on a B200)."""
    from generate3d_b200 import dfgen, _lib
    if tiled is None:
        monkeypatch.delenv("G3D_INIT_TILED", raising=False)
    else:
        monkeypatch.setenv("G3D_INIT_TILED", tiled)
    rng = np.random.default_rng(2024)
    ni, nj, nk = 20, 37, 45
    Vs, Ts, vo, to = [], [], [0], [0]
    for m in range(48):
        kind = m % 4
        if kind == 0:
            v, t = S.table_mesh(int(rng.integers(2, 9)), 500 + m, scale=rng.uniform(0.4, 1.0), rot_y=rng.uniform(0, 6.28))
        elif kind == 1:
            v = rng.uniform(-0.6, 0.6, (60, 3)).astype(np.float32)
            t = rng.integers(0, 60, (int(rng.integers(1, 200)), 3)).astype(np.uint32)
        elif kind == 2:                                            # a few triangles larger than the grid
            v = rng.uniform(-3, 3, (12, 3)).astype(np.float32)
            t = rng.integers(0, 12, (int(rng.integers(1, 6)), 3)).astype(np.uint32)
        else:                                                      # long thin triangles: wide in i, few (j, k) rows
            v = rng.uniform(-0.5, 0.5, (40, 3)).astype(np.float32)
            v[:, 1:] *= np.float32(0.05)
            t = rng.integers(0, 40, (int(rng.integers(1, 80)), 3)).astype(np.uint32)
        Vs.append(v); Ts.append(t)
        vo.append(vo[-1] + len(v)); to.append(to[-1] + len(t))
    V, T = np.concatenate(Vs), np.concatenate(Ts)
    p = _lib.TdfParams()
    p.ni, p.nj, p.nk = ni, nj, nk
    origin = np.array([-0.5, -0.55, -0.6], np.float32)
    dx = np.float32(1.0 / 36)
    p.origin[:] = [float(x) for x in origin]
    p.dx, p.out_scale, p.trunc, p.occ_thresh, p.mode, p.exact_band = float(dx), 1.0, float("inf"), 1.0, 0, 1
    r = dfgen.tdf_batch_params(V, np.array(vo), T, np.array(to), p, ("sdf", "closest_tri", "status"))
    assert not r["status"].cpu().numpy().any()
    sdf, ct = r["sdf"].cpu().numpy(), r["closest_tri"].cpu().numpy()
    for m in range(48):
        want = oracle.make_level_set3(Ts[m], Vs[m], origin, dx, ni, nj, nk)
        assert np.array_equal(_bits(sdf[m]), _bits(want["phi"])), m
        assert np.array_equal(ct[m], want["closest_tri"]), m


def test_host_abi_submit_wait_pipeline(g3d):
    """g3d_tdf_batch_host_submit / _wait: two batches in flight on one context give the results of the synchronous
    call, tickets are checked, a third submit is refused, raycast_host is refused while a batch is in flight."""
    from generate3d_b200 import dfgen
    L = g3d.lib()
    ctx = C.c_void_p()
    g3d.check(L.g3d_context_create(0, C.byref(ctx)))
    p = g3d.TdfParams()
    g2w = np.zeros(16, np.float32)
    g3d.check(L.g3d_dfgen_geometry(32, 0, 1, None, None, C.byref(p), g2w.ctypes.data))
    p.trunc, p.occ_thresh, p.mode = 3.0, 1.0, 0
    batches = []
    for b, (count, n, seed) in enumerate(((80, 3, 900), (70, 2, 901), (96, 2, 902))):
        V, voff, T, toff = S.table_batch(count, n, first_seed=seed)
        hV, hT = torch.from_numpy(V).pin_memory(), torch.from_numpy(T.view(np.int32)).pin_memory()
        sdf = torch.empty((count, 32, 32, 32), dtype=torch.float32).pin_memory()
        st = torch.empty((count,), dtype=torch.int32).pin_memory()
        o = g3d.TdfOutputs()
        o.sdf, o.status = sdf.data_ptr(), st.data_ptr()
        batches.append((V, voff, T, toff, hV, hT, sdf, st, o, count))

    def submit(bt):
        tk = C.c_int32(-1)
        rc = L.g3d_tdf_batch_host_submit(ctx, bt[4].data_ptr(), bt[1].ctypes.data, bt[5].data_ptr(), bt[3].ctypes.data,
                                         bt[9], C.byref(p), C.byref(bt[8]), C.byref(tk))
        return rc, tk.value
    rc0, t0 = submit(batches[0])
    rc1, t1 = submit(batches[1])
    assert rc0 == 0 and rc1 == 0 and {t0, t1} == {0, 1}
    rc2, t2 = submit(batches[2])
    assert rc2 == -1 and t2 == -1 and b"in flight" in L.g3d_last_error()      # two is the limit
    d = torch.zeros((1, 64, 64), dtype=torch.uint8)
    assert L.g3d_raycast_batch_host(ctx, d.data_ptr(), g2w.ctypes.data, 1, 64, 64, 65.0, 32.0, 32.0, 511, 32, None, None, None) == -1
    g3d.check(L.g3d_tdf_batch_host_wait(ctx, t0))
    assert L.g3d_tdf_batch_host_wait(ctx, t0) == -1                           # already collected
    rc2, t2 = submit(batches[2])                                              # the freed slot is reused while t1 is in flight
    assert rc2 == 0 and t2 == t0
    g3d.check(L.g3d_tdf_batch_host_wait(ctx, t1))
    g3d.check(L.g3d_tdf_batch_host_wait(ctx, t2))
    assert L.g3d_tdf_batch_host_wait(ctx, 5) == -1
    for bt in batches:
        want = dfgen.tdf_batch(bt[0], bt[1], bt[2], bt[3], outputs=("sdf",))["sdf"].cpu().numpy()
        assert not bt[7].numpy().any()
        assert np.array_equal(_bits(bt[6].numpy()), _bits(want))
    # an empty batch is a valid ticket too
    tk = C.c_int32(-1)
    g3d.check(L.g3d_tdf_batch_host_submit(ctx, None, None, None, None, 0, C.byref(p), C.byref(batches[0][8]), C.byref(tk)))
    g3d.check(L.g3d_tdf_batch_host_wait(ctx, tk.value))
    L.g3d_context_destroy(ctx)


# ---------------------------------------------------------------- BASELINE configs[0], [2], [3], [4] at full size

@pytest.fixture(scope="module")
def config3_case(oracle):
    """BASELINE configs[3]: one 201,840-triangle table at 128^3. The reference itself (oracle/_ref, else the C port)
    computes the expected field once per test session (~10 s of CPU)."""
    v, t = S.table_mesh(58, seed=0)
    g = oracle.dfgen_geometry(128)
    if oracle.ref_available():
        phi = oracle.ref_make_level_set3(t, v, g["origin"], g["dx"], 128, 128, 128)
    else:
        phi = oracle.make_level_set3(t, v, g["origin"], g["dx"], 128, 128, 128)["phi"]
    return v, t, phi * np.float32(128)


@pytest.mark.parametrize("cluster", [None, "2", "4", "8", "16", "8-plain", "4-spec"])
def test_config3_128cube_200k_tris_faithful_bit_equal(g3d, oracle, config3_case, cluster, monkeypatch):
    """configs[3] (128^3 x 201,840 triangles), faithful mode, bit-equal to the reference's make_level_set3 with the
    sweep's thread-block cluster size left to the library (16) and forced to 2, 4, 8 and 16 CTAs per mesh (second pass in the
    cluster form of the bitmap kernel, activity bits in global memory), with all 16 sweeps in the plain cluster kernel
    ("8-plain") and with speculation forced in every second-pass sweep ("4-spec")."""
    from generate3d_b200 import dfgen
    v, t, want = config3_case
    assert len(t) == 201840
    for k in ("G3D_SWEEP_CLUSTER", "G3D_SWEEP_BITMAP", "G3D_SWEEP_SPEC"):
        monkeypatch.delenv(k, raising=False)
    if cluster is not None:
        monkeypatch.setenv("G3D_SWEEP_CLUSTER", cluster.split("-")[0])
        if cluster.endswith("-plain"):
            monkeypatch.setenv("G3D_SWEEP_BITMAP", "0")
        if cluster.endswith("-spec"):
            monkeypatch.setenv("G3D_SWEEP_SPEC", "-1")
    got = _run(dfgen, [(v, t)], 128, outputs=("sdf", "occ_f32", "status"))
    assert got["status"][0] == 0
    assert np.array_equal(_bits(got["sdf"][0]), _bits(want))
    assert np.array_equal(got["occ_f32"][0].astype(np.uint8), oracle.occupancy(want))


def test_config3_128cube_exact_mode_properties(g3d, oracle, config3_case):
    """configs[3] in exact mode: a CPU brute force over 4.2e11 pairs is out of reach, so the size-independent
    properties are checked instead: same signs, identical |d| <= 1 occupancy (the reference's band is exact there),
    bit-equal distances inside that band, never larger than the reference's field within the truncation, and equal
    to the brute-force oracle on a sample of 1,000 voxels."""
    from generate3d_b200 import dfgen
    v, t, want = config3_case
    got = _run(dfgen, [(v, t)], 128, mode="exact", outputs=("sdf", "occ_f32", "closest_tri", "status"))
    sdf = got["sdf"][0]
    assert got["status"][0] == 0
    assert np.array_equal(np.signbit(sdf), np.signbit(want))
    assert np.array_equal(got["occ_f32"][0].astype(np.uint8), oracle.occupancy(want))
    band1 = np.abs(want) <= 1.0
    assert np.array_equal(_bits(sdf[band1]), _bits(want[band1]))
    near = np.signbit(want) | (np.abs(sdf) <= 3.0)
    assert np.all(np.abs(sdf)[near] <= np.abs(want)[near])
    # sampled brute force (the reference's distance function over ALL triangles) on voxels of the band
    rng = np.random.default_rng(7)
    idx = np.flatnonzero(np.abs(want).ravel() <= 3.0)
    pick = rng.choice(idx, 1000, replace=False)
    g = oracle.dfgen_geometry(128)
    bf, bt = oracle.brute_force_points(t, v, g["origin"], g["dx"], 128, 128, 128, pick)
    assert np.array_equal(_bits(np.abs(sdf).ravel()[pick]), _bits(bf * np.float32(128)))
    assert np.array_equal(got["closest_tri"][0].ravel()[pick], bt)


@pytest.mark.parametrize("mode", ["faithful", "exact"])
def test_config0_single_5k_mesh_both_modes(g3d, oracle, mode):
    """configs[0]: one 4,860-triangle table at 32^3, truncation 3: sdf, tdf, tdflog and occupancy against the oracle
    (faithful: everything bit-equal / 1e-5; exact: bit-equal to the brute force where it matters)."""
    from generate3d_b200 import dfgen
    v, t = S.table_mesh(9, seed=0)
    assert len(t) == 4860
    want = _oracle_full(oracle, v, t, 32)
    got = _run(dfgen, [(v, t)], 32, mode=mode)
    assert np.array_equal(got["occ_f32"][0].astype(np.uint8), oracle.occupancy(want["sdf"]))
    if mode == "faithful":
        assert np.array_equal(_bits(got["sdf"][0]), _bits(want["sdf"]))
        tdf = oracle.load_df_clamp(want["sdf"], 3.0)
        assert np.array_equal(_bits(got["tdf"][0]), _bits(tdf))
        assert np.abs(got["tdflog"][0] - oracle.apply_log_transform(tdf)).max() <= 1e-5
    else:
        g = oracle.dfgen_geometry(32)
        bf, bt = oracle.brute_force(t, v, g["origin"], g["dx"], 32, 32, 32)
        near = np.signbit(want["sdf"]) | (bf * np.float32(32) <= 3.0)
        assert np.array_equal(_bits(np.abs(got["sdf"][0])[near]), _bits((bf * np.float32(32))[near]))
        assert np.array_equal(got["closest_tri"][0][near], bt[near])
        assert np.abs(got["tdflog"][0][near] - oracle.apply_log_transform(oracle.load_df_clamp(
            np.copysign(bf * np.float32(32), want["sdf"]), 3.0))[near]).max() <= 1e-5


def test_config2_bench_batch_sampled_meshes_vs_reference(g3d, oracle):
    """configs[2] at FULL size: the very batch bench.py times (1,024 tables x 19,440 triangles, seeds 0..1023), all
    five outputs; 8 sampled meshes are compared bit for bit with the reference's make_level_set3, and the whole
    batch must be free of error flags and equal to itself when solved a second time (determinism)."""
    from generate3d_b200 import dfgen
    V, voff, T, toff = S.table_batch(1024, 18, first_seed=0)
    outs = ("sdf", "tdf", "tdflog", "occ_bits", "status")
    r = dfgen.tdf_batch(V, voff, T, toff, outputs=outs)
    torch.cuda.synchronize()
    assert int(r["status"].sum().item()) == 0
    sdf = r["sdf"].cpu().numpy()
    bits = r["occ_bits"].cpu().numpy()
    tdflog = r["tdflog"].cpu().numpy()
    g = oracle.dfgen_geometry(32)
    for m in (0, 1, 255, 256, 511, 700, 1022, 1023):
        v, t = V[voff[m]:voff[m + 1]], T[toff[m]:toff[m + 1]]
        if oracle.ref_available():
            want = oracle.ref_make_level_set3(t, v, g["origin"], g["dx"], 32, 32, 32) * np.float32(32)
        else:
            want = _oracle_full(oracle, v, t, 32)["sdf"]
        assert np.array_equal(_bits(sdf[m]), _bits(want)), m
        occ = np.unpackbits(bits[m].view(np.uint8), bitorder="little").reshape(32, 32, 32)
        assert np.array_equal(occ, oracle.occupancy(want)), m
        assert np.abs(tdflog[m] - oracle.apply_log_transform(oracle.load_df_clamp(want, 3.0))).max() <= 1e-5, m
    r2 = dfgen.tdf_batch(V, voff, T, toff, outputs=("sdf",))
    assert torch.equal(r2["sdf"], r["sdf"])


def test_config4_device_generated_shard_matches_host_path(g3d, oracle):
    """configs[4] (dataset-scale generation, inputs generated on the device per shard from the model seed): a chunk
    of the on-device generator, copied back, gives the same fields through the CPU reference; chunks are
    independent of chunk size and of the rank that owns them."""
    from generate3d_b200 import dataset_gen
    dev = torch.device("cuda", 0)
    a = dataset_gen.generate_tables(first_model=4000, count=6, n=18, device=dev)
    b = dataset_gen.generate_tables(first_model=4003, count=3, n=18, device=dev)
    nv = a["verts"].shape[0] // 6
    assert torch.equal(a["verts"][3 * nv:], b["verts"])                      # a model depends on its index only
    res = dataset_gen.voxelize_shard(first_model=4000, count=6, chunk=4, n=18, device=dev, keep=("sdf", "tdf", "occ_bits"))
    assert int(res["status"].sum().item()) == 0 and res["sdf"].shape[0] == 6
    V = a["verts"].cpu().numpy()
    T = a["tris"].cpu().numpy().view(np.uint32)
    g = oracle.dfgen_geometry(32)
    nt = T.shape[0] // 6
    for m in (0, 3, 5):
        want = _oracle_full(oracle, V[m * nv:(m + 1) * nv], T[m * nt:(m + 1) * nt], 32)["sdf"]
        assert np.array_equal(_bits(res["sdf"][m].cpu().numpy()), _bits(want)), m
    lo, hi = dataset_gen.shard_models(100000, 3, 8)
    assert (lo, hi) == (37500, 50000)


def test_two_streams_do_not_share_scratch(g3d):
    """the device ABI is thread-safe for distinct streams. Two batches enqueued back to back on two streams
    (so that their kernels may interleave) give the results of the same batches solved one after the other."""
    from generate3d_b200 import dfgen
    A = S.table_batch(40, 6, first_seed=700)
    B = S.table_batch(40, 6, first_seed=800)
    ra = dfgen.tdf_batch(*A, outputs=("sdf", "closest_tri"))
    rb = dfgen.tdf_batch(*B, outputs=("sdf", "closest_tri"))
    torch.cuda.synchronize()
    dA = [torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda() for x in A]
    dB = [torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda() for x in B]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(3):
        with torch.cuda.stream(s1):
            qa = dfgen.tdf_batch(*dA, outputs=("sdf", "closest_tri"))
        with torch.cuda.stream(s2):
            qb = dfgen.tdf_batch(*dB, outputs=("sdf", "closest_tri"))
    torch.cuda.synchronize()
    assert torch.equal(qa["sdf"], ra["sdf"]) and torch.equal(qa["closest_tri"], ra["closest_tri"])
    assert torch.equal(qb["sdf"], rb["sdf"]) and torch.equal(qb["closest_tri"], rb["closest_tri"])


def _write_obj(path, v, t, fmt="%.9g"):
    with open(path, "w") as f:
        for p in v:
            f.write(("v " + fmt + " " + fmt + " " + fmt + "\n") % tuple(p))
        for tri in t:
            f.write("f %d %d %d\n" % tuple(int(x) + 1 for x in tri))


def test_dfgen_cli_batch_mode_equals_reference_program(g3d, oracle, tmp_path):
    """`dfgen --batch list dim padding unit` (SURVEY 8b extension): one process and one CUDA context for a list of
    models. Every .df is byte-identical to what the reference program writes for that model (oracle/_ref/dfgen_ref:
    reference core + reference OBJ reader run as a program; the C port where _ref is absent), the .txt equals
    vox2mesh's rows, an unreadable model is reported (exit code 4) without disturbing the others, and chunking does
    not change a byte."""
    cli = os.path.join(os.path.dirname(g3d.LIB_PATH), "bin", "dfgen")
    ref_exe = os.path.join(os.path.dirname(oracle.__file__), "_ref", "dfgen_ref")
    rows = []
    for k in range(7):
        v, t = S.table_mesh(2 + k % 4, 80 + k, scale=0.75 + 0.03 * k, rot_y=0.4 * k)
        obj = tmp_path / ("m%d.obj" % k)
        _write_obj(obj, v, t, "%.7f" if k % 2 else "%.9g")
        rows.append((str(obj), str(tmp_path / ("m%d__0__.df" % k)), v, t))
    broken = tmp_path / "broken.obj"
    broken.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 0 1 2\n")
    lst = tmp_path / "list.txt"
    lst.write_text("".join("%s %s\n" % (r[0], r[1]) for r in rows[:4]) + "%s %s\n" % (broken, tmp_path / "broken.df") +
                   "".join("%s %s\n" % (r[0], r[1]) for r in rows[4:]))
    out = subprocess.run([cli, "--batch", str(lst), "32", "0", "1", "--occ-txt", "--chunk", "3"], capture_output=True, text=True)
    assert out.returncode == 4 and "7 of 8 models written, 1 failed" in out.stdout, out.stdout[-500:] + out.stderr[-500:]
    assert not (tmp_path / "broken.df").exists()
    for objp, dfp, v, t in rows:
        want = tmp_path / "want.df"
        if os.path.exists(ref_exe):
            subprocess.check_call([ref_exe, objp, "32", "0", "1", str(want)], stdout=subprocess.DEVNULL)
        else:
            v32, t32 = np.loadtxt(objp, usecols=(1, 2, 3), max_rows=len(v), dtype=np.float32), t
            g = oracle.dfgen(t32, v32, 32)
            oracle.write_df(str(want), g["sdf"], g["grid2world"])
        assert open(dfp, "rb").read() == want.read_bytes(), objp
        sdf = np.frombuffer(want.read_bytes()[80:], np.float32).reshape(32, 32, 32)
        assert open(dfp[:-3] + ".txt").read() == oracle.gt_occ_txt_rows(sdf)
    # one chunk for everything gives the same files; the single-model form too
    lst2 = tmp_path / "list2.txt"
    lst2.write_text("".join("%s %s\n" % (r[0], str(tmp_path / ("b%d.df" % k))) for k, r in enumerate(rows)))
    assert subprocess.call([cli, "--batch", str(lst2), "32", "0", "1"], stdout=subprocess.DEVNULL) == 0
    for k, r in enumerate(rows):
        assert (tmp_path / ("b%d.df" % k)).read_bytes() == open(r[1], "rb").read()
    single = tmp_path / "single.df"
    subprocess.check_call([cli, rows[2][0], "32", "0", "1", str(single)], stdout=subprocess.DEVNULL)
    assert single.read_bytes() == open(rows[2][1], "rb").read()
    # is_unit_cube = 0: every model on its own bounding box (one submit per model), padding 1
    lst3 = tmp_path / "list3.txt"
    lst3.write_text("".join("%s %s\n" % (r[0], str(tmp_path / ("u%d.df" % k))) for k, r in enumerate(rows[:3])))
    rc = subprocess.call([cli, "--batch", str(lst3), "32", "1", "0"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for k, r in enumerate(rows[:3]):
        one = tmp_path / "one.df"
        rc1 = subprocess.call([cli, r[0], "32", "1", "0", str(one)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if rc1 == 0:
            assert (tmp_path / ("u%d.df" % k)).read_bytes() == one.read_bytes()
    assert rc in (0, 4)


def test_batch_of_more_than_2pow27_triangles_in_total(g3d, oracle):
    """only the per-MESH triangle count is bounded by the 27-bit closest_tri field; a call may carry more
    than 2^27 triangles in total. 1,060 copies of a 126,960-triangle table (134.6 M triangles, built on the device)
    give 1,060 times the single-mesh field, which equals the oracle's."""
    from generate3d_b200 import dfgen
    v, t = S.table_mesh(46, 5)
    copies = 1060
    assert len(t) * copies > (1 << 27)
    dv = torch.from_numpy(v).cuda().repeat(copies, 1)
    dt = torch.from_numpy(t.view(np.int32)).cuda().repeat(copies, 1)
    voff = torch.arange(copies + 1, dtype=torch.int64, device="cuda") * len(v)
    toff = torch.arange(copies + 1, dtype=torch.int64, device="cuda") * len(t)
    r = dfgen.tdf_batch(dv, voff, dt, toff, outputs=("sdf", "closest_tri", "status"))
    torch.cuda.synchronize()
    assert int(r["status"].abs().sum().item()) == 0
    want = _oracle_full(oracle, v, t, 32)
    assert np.array_equal(_bits(r["sdf"][0].cpu().numpy()), _bits(want["sdf"]))
    assert np.array_equal(r["closest_tri"][copies - 1].cpu().numpy(), want["closest_tri"])
    assert bool((r["sdf"] == r["sdf"][0:1]).all()) and bool((r["closest_tri"] == r["closest_tri"][0:1]).all())
    del r, dv, dt
    torch.cuda.empty_cache()


def test_cadvoxelization_isolates_a_failing_batch_and_writes_ply(g3d, oracle, tmp_path, monkeypatch):
    """a g3d_tdf_batch call that fails as a whole must not take the chunk down: the chunk is solved again
    model by model and only the offender is reported. Also: write_ply=True leaves vox2mesh's .txt + .ply, and a
    triangle budget splits heavy chunks without changing a byte."""
    from generate3d_b200 import CADVoxelization as CV, dfgen, _lib
    root = tmp_path / "shapenet"
    want = {}
    for k in range(5):
        v, t = S.table_mesh(3, 160 + k, scale=0.9)
        d = root / "04379243" / ("model%02d" % k) / "models"
        d.mkdir(parents=True)
        _write_obj(d / "model_normalized.obj", v, t)
        want["model%02d" % k] = oracle.dfgen(t, v, 32)
    models = CV.list_models({"shapenet": str(root)})
    real = dfgen.tdf_batch
    calls = []

    def flaky(V, voff, T, toff, **kw):
        n = len(voff) - 1
        calls.append(n)
        if n > 1:
            raise _lib.G3DError(-2, "simulated device failure for a multi-model call")
        if len(calls) == 4:                                         # the third single-model retry
            raise _lib.G3DError(-1, "simulated bad model")
        return real(V, voff, T, toff, **kw)
    monkeypatch.setattr(dfgen, "tdf_batch", flaky)
    out_root = str(tmp_path / "vox") + "/"
    done, failed = CV.voxelize_models(models, out_root, chunk=8, write_ply=True)
    assert calls[0] == 5 and calls[1:] == [1] * 5
    assert done == 4 and len(failed) == 1 and "simulated bad model" in failed[0][1]
    bad = failed[0][0]
    for mid, g in want.items():
        out = tmp_path / "vox" / "04379243" / (mid + "__0__.df")
        if mid == bad:
            assert not out.exists()
            continue
        ref = tmp_path / "ref.df"
        oracle.write_df(str(ref), g["sdf"], g["grid2world"])
        assert out.read_bytes() == ref.read_bytes()
        assert (tmp_path / "vox" / "04379243" / (mid + "__0__.txt")).read_text() == oracle.gt_occ_txt_rows(g["sdf"])
        assert (tmp_path / "vox" / "04379243" / (mid + "__0__.ply")).read_bytes().startswith(b"ply\nformat ascii 1.0\n")
    monkeypatch.setattr(dfgen, "tdf_batch", real)
    calls.clear()
    done2, failed2 = CV.voxelize_models(models, str(tmp_path / "vox2") + "/", chunk=8, tri_budget=1100)   # 540 tris per model
    assert done2 == 5 and not failed2
    for mid in want:
        a = (tmp_path / "vox2" / "04379243" / (mid + "__0__.df")).read_bytes()
        ref = tmp_path / "ref.df"
        oracle.write_df(str(ref), want[mid]["sdf"], want[mid]["grid2world"])
        assert a == ref.read_bytes()


def test_host_abi_u16_indices_equal_u32(g3d):
    """g3d_tdf_batch_host_submit_u16: the same batch with 16-bit triangle indices gives the same bits; a mesh with more
    than 65,536 vertices is refused with a message."""
    V, voff, T, toff = S.table_batch(70, 3, first_seed=400)
    L = g3d.lib()
    ctx = C.c_void_p()
    g3d.check(L.g3d_context_create(0, C.byref(ctx)))
    p = g3d.TdfParams()
    g2w = np.zeros(16, np.float32)
    g3d.check(L.g3d_dfgen_geometry(32, 0, 1, None, None, C.byref(p), g2w.ctypes.data))
    p.trunc, p.occ_thresh, p.mode = 3.0, 1.0, 0
    res = []
    for bits in (32, 16):
        sdf = np.empty((70, 32, 32, 32), np.float32)
        ct = np.empty((70, 32, 32, 32), np.int32)
        o = g3d.TdfOutputs()
        o.sdf, o.closest_tri = sdf.ctypes.data, ct.ctypes.data
        tk = C.c_int32(-1)
        if bits == 32:
            g3d.check(L.g3d_tdf_batch_host_submit(ctx, V.ctypes.data, voff.ctypes.data, T.ctypes.data, toff.ctypes.data, 70, C.byref(p), C.byref(o), C.byref(tk)))
        else:
            T16 = np.ascontiguousarray(T.astype(np.uint16))
            g3d.check(L.g3d_tdf_batch_host_submit_u16(ctx, V.ctypes.data, voff.ctypes.data, T16.ctypes.data, toff.ctypes.data, 70, C.byref(p), C.byref(o), C.byref(tk)))
        g3d.check(L.g3d_tdf_batch_host_wait(ctx, tk.value))
        res.append((sdf, ct))
    assert np.array_equal(_bits(res[0][0]), _bits(res[1][0])) and np.array_equal(res[0][1], res[1][1])
    big_off = np.array([0, 70000], np.int64)
    tk = C.c_int32(-1)
    o = g3d.TdfOutputs()
    rc = L.g3d_tdf_batch_host_submit_u16(ctx, V.ctypes.data, big_off.ctypes.data, T.ctypes.data, np.array([0, 1], np.int64).ctypes.data, 1,
                                         C.byref(p), C.byref(o), C.byref(tk))
    assert rc == -1 and b"65,536" in L.g3d_last_error()
    L.g3d_context_destroy(ctx)

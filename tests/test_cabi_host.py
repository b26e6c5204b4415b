"""CPU tests of the product library's host-side surface: it loads, exports every symbol
include/g3d.h declares, and its non-GPU entry points (driver arithmetic, .df and mesh I/O)
agree with the oracle. No compute entry point is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from generate3d_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(g3d):
    L = g3d.lib()
    header = open(os.path.join(ROOT, "include", "g3d.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(g3d_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 18
    for n in sorted(names):
        assert hasattr(L, n), n
    assert names == set(g3d.SYMBOLS), names ^ set(g3d.SYMBOLS)
    assert L.g3d_version() >= 100


def test_no_gpu_means_error_not_fallback(g3d):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from generate3d_b200 import dfgen
    v, t = S.table_mesh(2, 0)
    with pytest.raises(RuntimeError):
        dfgen.tdf_batch(v, np.array([0, len(v)]), t, np.array([0, len(t)]))
    ctx = C.c_void_p()
    assert g3d.lib().g3d_context_create(0, C.byref(ctx)) != 0
    assert g3d.lib().g3d_last_error()
    # the projection-layer mirror and the split host call have no CPU path either
    from generate3d_b200 import perspective_projection as pp
    with pytest.raises(RuntimeError):
        pp.ProjectionHelper()
    tk = C.c_int32(7)
    assert g3d.lib().g3d_tdf_batch_host_submit(None, None, None, None, None, 0, None, None, C.byref(tk)) == -1
    assert g3d.lib().g3d_tdf_batch_host_wait(None, 0) == -1


@pytest.mark.parametrize("dim,pad,unit", [(32, 0, 1), (64, 0, 1), (128, 0, 1), (32, 1, 1), (32, 2, 1), (20, 0, 1)])
def test_dfgen_geometry_matches_oracle(g3d, oracle, dim, pad, unit):
    from generate3d_b200 import dfgen
    want = oracle.dfgen_geometry(dim, pad, unit)
    try:
        p, g2w = dfgen.dfgen_geometry(dim, pad, unit)
        ok = True
    except g3d.G3DError as e:
        ok = False
        assert e.code == -5
    assert ok == want["ok"]
    if ok:
        assert (p.ni, p.nj, p.nk) == tuple(int(s) for s in want["sizes"])
        assert np.float32(p.dx) == want["dx"] and p.out_scale == dim
        assert np.array_equal(np.float32(list(p.origin)), want["origin"])
        assert np.array_equal(g2w, want["grid2world"])


def test_df_roundtrip_and_golden_bytes(g3d, golden_dir, tmp_path):
    from generate3d_b200 import dfgen
    z = np.load(os.path.join(golden_dir, "loader.npz"))
    src = tmp_path / "g.df"
    src.write_bytes(bytes(z["df_file"]))
    sdf, res, g2w = dfgen.read_df(str(src))
    assert sdf.shape == (32, 32, 32) and res == 1.0
    out = tmp_path / "o.df"
    dfgen.write_df(str(out), sdf, g2w)
    assert out.read_bytes() == bytes(z["df_file"])


def test_load_mesh_obj_and_ply(g3d, tmp_path):
    from generate3d_b200 import dfgen
    v, t = S.table_mesh(2, 1)
    obj = tmp_path / "m.obj"
    with open(obj, "w") as f:
        f.write("# comment\nmtllib x.mtl\n")
        for p in v:
            f.write("v %.9g %.9g %.9g\n" % tuple(p))
        f.write("v 9 9 9\n")                                  # unreferenced vertex is kept
        f.write("g part\n")
        for k, tri in enumerate(t):
            if k % 3 == 0:
                f.write("f %d %d %d\n" % tuple(tri + 1))
            elif k % 3 == 1:
                f.write("f %d/1/1 %d/2/2 %d/3/3\n" % tuple(tri + 1))
            else:
                f.write("f %d %d %d\n" % tuple(tri.astype(np.int64) - (len(v) + 1)))   # relative
        f.write("f 1 2 3 4\n")                                # quad -> fan
    v2, t2 = dfgen.load_mesh(str(obj))
    assert v2.shape == (len(v) + 1, 3) and np.array_equal(v2[:-1], v)
    assert np.array_equal(t2[:len(t)], t)
    assert t2[len(t):].tolist() == [[0, 1, 2], [0, 2, 3]]
    ply = tmp_path / "m.ply"
    with open(ply, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
                 "property float z\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n"
                 % (len(v), len(t))).encode())
        f.write(v.tobytes())
        for tri in t:
            f.write(b"\x03" + tri.astype(np.int32).tobytes())
    v3, t3 = dfgen.load_mesh(str(ply))
    assert np.array_equal(v3, v) and np.array_equal(t3, t)
    with pytest.raises(g3d.G3DError):
        dfgen.load_mesh(str(tmp_path / "m.stl"))


def test_read_df_rejects_overflowing_headers(g3d, tmp_path):
    """Header dims are untrusted: three large dims whose product wraps must not slip past the capacity check."""
    from generate3d_b200 import dfgen
    L = g3d.lib()
    bad = tmp_path / "bad.df"
    for dims in ([1 << 21, 1 << 21, 1 << 21], [70000, 1, 1], [-1, 4, 4], [65536, 65536, 65536]):
        with open(bad, "wb") as f:
            f.write(np.asarray(dims, np.int32).tobytes() + np.float32(1).tobytes() + np.zeros(16, np.float32).tobytes())
            f.write(np.zeros(4096, np.float32).tobytes())
        out = np.zeros(64, np.float32)
        d3 = np.zeros(3, np.int32)
        rc = L.g3d_read_df(os.fsencode(str(bad)), d3.ctypes.data, None, None, out.ctypes.data, 64)
        assert rc == -1 and not out.any(), dims
    with pytest.raises(g3d.G3DError):
        dfgen.read_df(str(bad))


def _vox2mesh_expected_arrays(sdf, g2w, res, trunc, unitless):
    """numpy restatement of get_position_and_color_from_vox (vox2mesh/main.cpp:37-84) for pdf = 0 / jet."""
    bv = np.float32([[-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1], [-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1]])
    be = np.uint32([[0, 1, 2], [2, 3, 0], [1, 5, 6], [6, 2, 1], [7, 6, 5], [5, 4, 7], [4, 0, 3], [3, 7, 4], [4, 5, 1], [1, 0, 4],
                    [3, 2, 6], [6, 7, 3]])
    M = g2w.reshape(4, 4).astype(np.float32)
    vs = np.sqrt((M[:3, :3] ** 2).sum(0, dtype=np.float32)).astype(np.float32) if unitless else np.ones(3, np.float32)
    r3 = (np.float32(0.45) * vs * np.float32(res)).astype(np.float32)
    k, j, i = np.nonzero(np.abs(sdf) <= np.float32(trunc) * np.float32(res))
    P = np.stack([((M[a, 0] * i.astype(np.float32) + M[a, 1] * j.astype(np.float32)) + M[a, 2] * k.astype(np.float32)) + M[a, 3]
                  for a in range(3)], 1).astype(np.float32)
    V = (bv[None] * r3[None, None] + P[:, None]).astype(np.float32).reshape(-1, 3)
    Cc = np.tile(np.uint8([0, 169, 255]), (len(V), 1))
    F = (be[None] + (np.arange(len(P), dtype=np.uint32) * 8)[:, None, None]).reshape(-1, 3).astype(np.uint32)
    return np.ascontiguousarray(V), np.ascontiguousarray(Cc), np.ascontiguousarray(F), (i, j, k)


@pytest.mark.parametrize("unitless", [0, 1])
def test_vox2mesh_txt_and_ply(g3d, oracle, tmp_path, unitless):
    """g3d_vox2mesh / bin/vox2mesh: the .txt equals the oracle's restatement of write_txt (vox2mesh/main.cpp:86-104);
    the PLY equals what the REFERENCE's tinyply writes (compiled in place) for the arrays main.cpp:37-84 builds."""
    import subprocess
    from generate3d_b200 import dfgen
    v, t = S.table_mesh(3, 5)
    g = oracle.dfgen(t, v, 32)
    df = tmp_path / "m__0__.df"
    oracle.write_df(str(df), g["sdf"], g["grid2world"])
    ply = tmp_path / "m__0__.ply"
    dfgen.vox2mesh(str(df), str(ply), is_unitless=bool(unitless))
    assert (tmp_path / "m__0__.txt").read_text() == oracle.gt_occ_txt_rows(g["sdf"])
    # the executable with the argv CADVoxelization.py:35 passes writes the same two files
    exe = os.path.join(os.path.dirname(g3d.LIB_PATH), "bin", "vox2mesh")
    ply2 = tmp_path / "cli.ply"
    subprocess.check_call([exe, "--in", str(df), "--out", str(ply2), "--is_unitless", str(unitless)])
    assert ply2.read_bytes() == ply.read_bytes()
    assert (tmp_path / "cli.txt").read_text() == (tmp_path / "m__0__.txt").read_text()
    assert subprocess.call([exe, "--in", str(df)], stderr=subprocess.DEVNULL) == 1          # --out is required
    assert subprocess.call([exe, "--in=" + str(df), "--out=" + str(ply2), "--trunc=0.5"]) == 0
    head = ply.read_bytes()[:400].decode()
    assert head.startswith("ply\nformat ascii 1.0\ncomment generated by tinyply 2.2\nelement vertex ")
    if oracle.ref_obj_available():
        V, Cc, F, _ = _vox2mesh_expected_arrays(g["sdf"], g["grid2world"], 1.0, 1.0, unitless)
        L = C.CDLL(os.path.join(os.path.dirname(oracle.__file__), "_ref", "libg3d_ref_obj.so"))
        want = tmp_path / "ref.ply"
        assert L.g3d_ref_write_ply(os.fsencode(str(want)), V.ctypes.data_as(C.c_void_p), C.c_int64(len(V)),
                                   Cc.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p), C.c_int64(len(F))) == 0
        assert ply.read_bytes() == want.read_bytes()
    with pytest.raises(g3d.G3DError):
        dfgen.vox2mesh(str(df), str(tmp_path / "noext"))
    with pytest.raises(g3d.G3DError):
        dfgen.vox2mesh(str(df), str(ply), cmap="magma")


def test_vox2mesh_pdf_block_and_redcenter(g3d, oracle, tmp_path):
    """A .df with a trailing pdf block (LoaderVOX.h:38-41) colours each visible voxel with jet(pdf) (Colormap.h:1301-1325,
    T = float with the header's double-typed literals); --redcenter marks the central 3x3x3 block when there is no pdf
    block (vox2mesh/main.cpp:180-192). Checked against a numpy restatement (Colormap.h needs Eigen: unpinned)."""
    from generate3d_b200 import dfgen
    rng = np.random.default_rng(9)
    dim = 8
    sdf = rng.uniform(-3, 3, (dim, dim, dim)).astype(np.float32)
    pdf = rng.uniform(-0.5, 1.5, (dim, dim, dim)).astype(np.float32)
    g2w = np.eye(4, dtype=np.float32).reshape(-1)

    def jet(v):
        v = np.float32(v)
        vmin, vmax = np.float32(-0.2), np.float32(1.0)
        v = min(max(v, vmin), vmax)
        dv = np.float32(vmax - vmin)
        c = [np.float32(1), np.float32(1), np.float32(1)]
        if float(v) < float(vmin) + 0.25 * float(dv):
            c[0] = np.float32(0)
            c[1] = np.float32(np.float32(np.float32(4) * np.float32(v - vmin)) / dv)
        elif float(v) < float(vmin) + 0.5 * float(dv):
            c[0] = np.float32(0)
            c[2] = np.float32(1 + 4 * (float(vmin) + 0.25 * float(dv) - float(v)) / float(dv))
        elif float(v) < float(vmin) + 0.75 * float(dv):
            c[0] = np.float32(4 * (float(v) - float(vmin) - 0.5 * float(dv)) / float(dv))
            c[2] = np.float32(0)
        else:
            c[1] = np.float32(1 + 4 * (float(vmin) + 0.75 * float(dv) - float(v)) / float(dv))
            c[2] = np.float32(0)
        return [int(np.float32(255.0) * x) for x in c]

    df = tmp_path / "p.df"
    with open(df, "wb") as f:
        f.write(np.int32([dim, dim, dim]).tobytes() + np.float32(1).tobytes() + g2w.tobytes() + sdf.tobytes() + pdf.tobytes())
    dfgen.vox2mesh(str(df), str(tmp_path / "p.ply"), write_ply=False)
    rows = (tmp_path / "p.txt").read_text().splitlines()
    want = []
    for k in range(dim):
        for j in range(dim):
            for i in range(dim):
                if abs(sdf[k, j, i]) <= 1.0:
                    want.append("%d %d %d %d %d %d" % (i, j, k, *jet(pdf[k, j, i])))
    assert rows == want and len(rows) > 50
    # no pdf block + --redcenter: pdf = 1 in the central block (index i*dim*dim + j*dim + k, as the reference writes it)
    with open(df, "wb") as f:
        f.write(np.int32([dim, dim, dim]).tobytes() + np.float32(1).tobytes() + g2w.tobytes() + np.zeros_like(sdf).tobytes())
    dfgen.vox2mesh(str(df), str(tmp_path / "p.ply"), redcenter=True, write_ply=False)
    rows = (tmp_path / "p.txt").read_text().splitlines()
    assert len(rows) == dim ** 3
    red = [r for r in rows if r.endswith(" %d %d %d" % tuple(jet(1.0)))]
    assert len(red) == 27 and rows[0].endswith(" 0 169 255")

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

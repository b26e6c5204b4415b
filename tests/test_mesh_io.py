"""g3d_load_mesh against the reference's OBJ reader (host logic, no GPU needed).

Pinned two ways: (1) tests/golden/obj_parse.npz -- an OBJ + MTL exercising number spellings, index forms, polygons,
groups / objects / materials and line endings, parsed by the reference's vendored tiny_obj_loader.h (compiled in
place by oracle/Makefile) through the gather loop of LoaderMesh.h:17-50; (2) when oracle/_ref is present (the build
container and the GPU box), live comparisons on random coordinates and random polygons.
"""
import os

import numpy as np
import pytest


def _write(path, text):
    with open(path, "w", newline="") as f:
        f.write(text)


def _tris(corners):
    return corners.astype(np.int64).astype(np.uint32).reshape(-1, 3)


def test_obj_golden_fixture(g3d, golden_dir, tmp_path):
    from generate3d_b200 import dfgen
    z = np.load(os.path.join(golden_dir, "obj_parse.npz"))
    _write(tmp_path / "model.obj", bytes(z["obj"]).decode())
    _write(tmp_path / "model.mtl", bytes(z["mtl"]).decode())
    v, t = dfgen.load_mesh(str(tmp_path / "model.obj"))
    assert np.array_equal(v.view(np.uint32), z["verts"].view(np.uint32))          # every coordinate, bit for bit
    assert np.array_equal(t, _tris(z["corners"]))                                # every emitted corner, in order
    for k, bad in enumerate(bytes(z["bad_obj"]).decode().split("\0")):           # inputs tinyobj refuses as a whole
        _write(tmp_path / ("bad%d.obj" % k), bad)
        with pytest.raises(g3d.G3DError):
            dfgen.load_mesh(str(tmp_path / ("bad%d.obj" % k)))


def test_obj_numbers_match_tinyobj_on_a_million_coordinates(g3d, oracle, tmp_path):
    """tiny_obj_loader.h:505-628 (digit accumulation in double, pow_lut / pow(10,-k) fractions, ldexp(m*5^e, e), then a
    cast to float) replayed by g3d_load_mesh: 1,020,000 coordinates with 4 to 17 significant digits, plain / exponent
    / repr forms, magnitudes 1e-6 .. 1e3, bit for bit."""
    if not oracle.ref_obj_available():
        pytest.skip("oracle/_ref/libg3d_ref_obj.so not built (needs /root/reference)")
    from generate3d_b200 import dfgen
    rng = np.random.default_rng(1)
    n = 340000
    vals = rng.uniform(-1, 1, (n, 3)) * 10.0 ** rng.integers(-6, 3, (n, 3))
    fmts = ("v %.6f %.6f %.6f\n", "v %.9g %.9g %.9g\n", "v %.17g %.17g %.17g\n", "v %.8e %.3E %+.12e\n", "v %r %r %r\n",
            "v %.4f %.1f %.12f\n")
    with open(tmp_path / "big.obj", "w") as f:
        f.write("".join(fmts[i % 6] % tuple(float(x) for x in vals[i]) for i in range(n)))
        f.write("f 1 2 3\n")
    rv, rf = oracle.ref_load_obj(str(tmp_path / "big.obj"))
    v, t = dfgen.load_mesh(str(tmp_path / "big.obj"))
    assert rv.shape == (n, 3) and np.array_equal(v.view(np.uint32), rv.view(np.uint32))
    assert np.array_equal(t, _tris(rf))


def test_obj_polygons_match_tinyobj_ear_clipping(g3d, oracle, tmp_path):
    """Random planar and non-planar polygons (convex, concave, self-touching, with repeated and collinear corners),
    4 to 12 corners, v/vt/vn index forms: the triangles tinyobj's ear clipping emits (tiny_obj_loader.h:1011-1146),
    in the same order."""
    if not oracle.ref_obj_available():
        pytest.skip("oracle/_ref/libg3d_ref_obj.so not built (needs /root/reference)")
    from generate3d_b200 import dfgen
    rng = np.random.default_rng(2)
    lines, nv = [], 0
    for case in range(3000):
        k = int(rng.integers(4, 13))
        kind = case % 5
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        rad = rng.uniform(0.2, 1.0, k) if kind in (1, 3) else np.full(k, rng.uniform(0.2, 1.0))       # star-shaped / convex
        P = np.stack([rad * np.cos(ang), rad * np.sin(ang), np.zeros(k)], 1)
        if kind == 2:
            P = rng.uniform(-1, 1, (k, 3))                                                             # arbitrary soup
        if kind == 3:
            P[:, 2] = rng.uniform(-0.05, 0.05, k)                                                      # nearly planar
        if kind == 4:
            P[rng.integers(k)] = P[rng.integers(k)]                                                    # a repeated corner
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        P = P @ R.T * rng.uniform(0.01, 2.0) + rng.uniform(-0.3, 0.3, 3)
        for p in P:
            lines.append("v %.7g %.7g %.7g" % tuple(p))
        idx = list(range(nv + 1, nv + k + 1))
        if case % 7 == 0:
            idx = idx[::-1]
        form = case % 4
        if form == 0:
            lines.append("f " + " ".join(str(i) for i in idx))
        elif form == 1:
            lines.append("f " + " ".join("%d/%d" % (i, 1) for i in idx))
        elif form == 2:
            lines.append("f " + " ".join("%d//%d" % (i, 1) for i in idx))
        else:
            lines.append("f " + " ".join(str(i - (nv + k) - 1) for i in idx))                          # relative indices
        nv += k
        if case % 97 == 0:
            lines.append("g part%d" % case)
    _write(tmp_path / "poly.obj", "vt 0 0\nvn 0 0 1\n" + "\n".join(lines) + "\n")
    rv, rf = oracle.ref_load_obj(str(tmp_path / "poly.obj"))
    v, t = dfgen.load_mesh(str(tmp_path / "poly.obj"))
    assert np.array_equal(v.view(np.uint32), rv.view(np.uint32))
    assert len(rf) % 3 == 0 and len(rf) // 3 > 3000
    assert np.array_equal(t, _tris(rf))


def test_obj_shape_bookkeeping_matches_tinyobj(g3d, oracle, tmp_path):
    """Random interleavings of f / g / o / usemtl lines with and without a material library: the surviving faces and
    their order (including the shape that version drops at an `o` line after a material change)."""
    if not oracle.ref_obj_available():
        pytest.skip("oracle/_ref/libg3d_ref_obj.so not built (needs /root/reference)")
    from generate3d_b200 import dfgen
    rng = np.random.default_rng(3)
    _write(tmp_path / "m.mtl", "newmtl a\nKd 1 1 1\nnewmtl b\nnewmtl a\n\nnewmtl c \n")
    for trial in range(60):
        lines = ["mtllib m.mtl"] if trial % 3 else (["mtllib nothere.mtl"] if trial % 2 else [])
        lines += ["v %d %d %d" % tuple(rng.integers(0, 5, 3)) for _ in range(12)]
        for _ in range(int(rng.integers(5, 40))):
            r = rng.integers(10)
            if r < 5:
                lines.append("f %d %d %d" % tuple(rng.integers(1, 13, 3)))
            elif r < 6:
                lines.append("g grp" if rng.integers(2) else "g")
            elif r < 7:
                lines.append("o thing")
            else:
                lines.append("usemtl " + ("a", "b", "c", "zzz", "a ")[int(rng.integers(5))])
        eol = ("\n", "\r\n", "\r")[trial % 3]
        _write(tmp_path / "s.obj", eol.join(lines) + (eol if trial % 2 else ""))
        ref = oracle.ref_load_obj(str(tmp_path / "s.obj"))
        v, t = dfgen.load_mesh(str(tmp_path / "s.obj"))
        assert np.array_equal(v, ref[0]), trial
        assert np.array_equal(t, _tris(ref[1])), (trial, lines)

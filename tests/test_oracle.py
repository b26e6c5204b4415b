"""CPU tests: the oracle against the compiled reference and the golden fixtures."""
import os

import numpy as np
import pytest

from generate3d_b200 import synthetic as S


def test_ptd_oracle_matches_compiled_reference(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    rng = np.random.default_rng(0)
    for i in range(5000):
        p = rng.uniform(-0.6, 0.6, (4, 3)).astype(np.float32)
        if i % 7 == 0:
            p[2] = p[1]                                   # zero-length edge -> NaN path
        if i % 11 == 0:
            p[3] = p[1] + (p[2] - p[1]) * np.float32(0.5)  # collinear
        if i % 13 == 0:
            p[1] = p[2] = p[3]                            # point triangle
        a = oracle.point_triangle_distance(*p)
        b = oracle.point_triangle_distance(*p, use_ref=True)
        assert a == b or (np.isnan(a) and np.isnan(b))


@pytest.mark.parametrize("n,dim", [(9, 32), (4, 16), (6, 40)])
def test_make_level_set3_oracle_bit_equal_to_reference(oracle, n, dim):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    v, t = S.table_mesh(n, seed=5)
    g = oracle.dfgen_geometry(dim)
    assert g["ok"] and tuple(g["sizes"]) == (dim, dim, dim)
    mine = oracle.make_level_set3(t, v, g["origin"], g["dx"], dim, dim, dim)["phi"]
    ref = oracle.ref_make_level_set3(t, v, g["origin"], g["dx"], dim, dim, dim)
    assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
    assert (ref < 0).any() and (ref > 0).any()


def test_make_level_set3_non_cubic(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    v, t = S.table_mesh(5, seed=2)
    origin = np.array([-0.55, -0.5, -0.45], np.float32)
    mine = oracle.make_level_set3(t, v, origin, 0.04, 27, 25, 23)["phi"]
    ref = oracle.ref_make_level_set3(t, v, origin, 0.04, 27, 25, 23)
    assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))


def test_dfgen_geometry_matches_reference_driver(oracle):
    g = oracle.dfgen_geometry(32, 0, 1)
    assert g["ok"] and g["dx"] == np.float32(0.03125)
    assert np.array_equal(g["origin"], np.float32([-0.5, -0.5, -0.5]))
    exp = np.float32([0.03125, 0, 0, -0.5, 0, 0.03125, 0, -0.5, 0, 0, 0.03125, -0.5, 0, 0, 0, 1])
    assert np.array_equal(g["grid2world"], exp)


def _check_raycast_fixture(oracle, path, W):
    z = np.load(path)
    focal = float(z["focal"])
    for q in range(z["pose"].shape[0]):
        im = oracle.compute_index_mapping(z["pose"][q], 32, focal, W / 2, W / 2, 511)
        assert np.array_equal(im, z["index_map"][q].astype(np.int64))
        occ = oracle.raycast_depth(z["depth"][q], z["pose"][q], 32, focal, W / 2, W / 2, 511)
        assert np.array_equal(np.packbits(occ.reshape(-1), bitorder="little"), z["occ_bits"][q])
    occ0 = np.unpackbits(z["occ_bits"][0], bitorder="little").reshape(32, 32, 32)
    assert oracle.occ_txt_rows(occ0) == bytes(z["first_txt"]).decode()


def test_raycast_oracle_vs_reference_512(oracle, golden_dir):
    _check_raycast_fixture(oracle, os.path.join(golden_dir, "raycast_512.npz"), 512)


def test_raycast_oracle_vs_reference_256(oracle, golden_dir):
    _check_raycast_fixture(oracle, os.path.join(golden_dir, "raycast_256.npz"), 256)


def test_raycast_oracle_vs_reference_random_poses(oracle, golden_dir):
    _check_raycast_fixture(oracle, os.path.join(golden_dir, "raycast_rand.npz"), 512)


def test_loader_golden(oracle, golden_dir, tmp_path):
    z = np.load(os.path.join(golden_dir, "loader.npz"))
    g = oracle.dfgen(z["tris"], z["verts"], 32)
    p = tmp_path / "m.df"
    oracle.write_df(str(p), g["sdf"], g["grid2world"])
    assert p.read_bytes() == bytes(z["df_file"])                      # byte-identical .df
    assert np.array_equal(oracle.load_df_clamp(g["sdf"], 3.0), z["tdf3"][0])
    assert np.array_equal(oracle.occupancy(g["sdf"]), z["occ_gt"][0])
    assert np.abs(oracle.apply_log_transform(z["tdf3"][0]) - z["tdflog3"][0]).max() <= 1e-6
    assert np.array_equal(S.view_pose(3), z["pose3"])


def test_synthetic_sizes():
    for n, T in ((9, 4860), (18, 19440)):
        v, t = S.table_mesh(n, 0)
        assert t.shape == (T, 3) and v.shape == (30 * (n + 1) ** 2, 3)
        assert np.abs(v).max() < 0.5 and t.max() < v.shape[0]
    d, p = S.depth_views(2, 64, 64)
    assert d.shape == (2, 64, 64) and d.dtype == np.uint8 and (d > 0).any() and p.shape == (2, 4, 4)


def test_reference_program_standin_reproduces_golden_df_bytes(oracle, golden_dir, tmp_path):
    """oracle/_ref/dfgen_ref (the reference's own makelevelset3.cpp + tiny_obj_loader.h compiled in place, main.cpp's
    Eigen-carrying glue restated) run as a PROGRAM on an OBJ file writes the .df bytes the fixture holds -- the same
    bytes the reference's load_df parsed when tests/golden/loader.npz was made -- and agrees with the C port's
    driver arithmetic for a padded grid."""
    import subprocess
    exe = os.path.join(os.path.dirname(oracle.__file__), "_ref", "dfgen_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dfgen_ref not built (needs /root/reference)")
    z = np.load(os.path.join(golden_dir, "loader.npz"))
    obj = tmp_path / "m.obj"
    with open(obj, "w") as f:
        for p in z["verts"]:
            f.write("v %.9g %.9g %.9g\n" % tuple(p))
        for tri in z["tris"]:
            f.write("f %d %d %d\n" % tuple(tri + 1))
    out = tmp_path / "o.df"
    subprocess.check_call([exe, str(obj), "32", "0", "1", str(out)])
    assert out.read_bytes() == bytes(z["df_file"])
    subprocess.check_call([exe, str(obj), "34", "1", "1", str(out)])
    g = oracle.dfgen(z["tris"], z["verts"], 34, 1, 1)
    ref = tmp_path / "r.df"
    oracle.write_df(str(ref), g["sdf"], g["grid2world"])
    assert out.read_bytes() == ref.read_bytes()
    assert subprocess.call([exe, str(obj), "32"], stdout=subprocess.DEVNULL) == 255


def test_raycast_loop_port_equals_vectorised_port_and_fixture(oracle, golden_dir):
    """The cost-faithful port of raycast_depth used as bench.py's CPU baseline (32,768-iteration Python loop over torch
    slices) returns exactly what the vectorised restatement and the reference's own output (fixture) hold."""
    z = np.load(os.path.join(golden_dir, "raycast_512.npz"))
    depth, pose = z["depth"][0], z["pose"][0]
    got = oracle.raycast_depth_reference_loop(depth, pose, 525.0)
    assert np.array_equal(got, oracle.raycast_depth(depth, pose, 32, 525.0))
    want = np.unpackbits(z["occ_bits"][0], bitorder="little").reshape(32, 32, 32)
    assert np.array_equal(got, want)

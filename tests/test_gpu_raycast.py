"""GPU parity tests of the raycast path against the reference's outputs (golden fixtures) and the oracle."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from generate3d_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _check_fixture(path, W):
    from generate3d_b200 import raytracing
    z = np.load(path)
    r = raytracing.raycast_batch(z["depth"], z["pose"], focal=float(z["focal"]), cx=W / 2, cy=W / 2,
                                 outputs=("occ_bits", "occ_f32", "index_map"))
    torch.cuda.synchronize()
    assert np.array_equal(r["index_map"].cpu().numpy(), z["index_map"].astype(np.int64))
    assert np.array_equal(r["occ_bits"].cpu().numpy().view(np.uint8), z["occ_bits"])
    occ = np.unpackbits(z["occ_bits"], axis=1, bitorder="little").reshape(-1, 32, 32, 32)
    assert np.array_equal(r["occ_f32"].cpu().numpy().astype(np.uint8), occ)
    return z, r


def test_raycast_vs_reference_512(g3d, golden_dir):
    z, r = _check_fixture(os.path.join(golden_dir, "raycast_512.npz"), 512)
    from generate3d_b200 import raytracing
    assert raytracing.occ_rows(r["occ_f32"][0]) == bytes(z["first_txt"]).decode()


def test_raycast_vs_reference_256_config1(g3d, golden_dir):
    _check_fixture(os.path.join(golden_dir, "raycast_256.npz"), 256)


def test_raycast_vs_reference_random_poses(g3d, golden_dir):
    _check_fixture(os.path.join(golden_dir, "raycast_rand.npz"), 512)


def test_raycast_file_dropin(g3d, golden_dir, tmp_path):
    """raycast_depth(txt, pose.json, depth.png) writes the reference's txt byte for byte."""
    from PIL import Image
    from generate3d_b200 import raytracing
    z = np.load(os.path.join(golden_dir, "raycast_512.npz"))
    Image.fromarray(z["depth"][0], mode="L").save(tmp_path / "depth00.png")
    with open(tmp_path / "pose00.json", "w") as f:
        json.dump({"pose": S.pose_json_payload(0)}, f)
    raytracing.raycast_depth(str(tmp_path / "o.txt"), str(tmp_path / "pose00.json"), str(tmp_path / "depth00.png"))
    assert (tmp_path / "o.txt").read_bytes() == bytes(z["first_txt"])
    im = raytracing.compute_index_mapping(raytracing.load_pose(str(tmp_path / "pose00.json")),
                                          torch.inverse(raytracing.make_world_to_grid()), torch.device("cuda"))
    assert im.dtype == torch.int64 and tuple(im.shape) == (4 * 32768,)
    assert np.array_equal(im.cpu().numpy().reshape(4, -1), z["index_map"][0].astype(np.int64))


def test_raycast_vs_oracle_many_poses_and_odd_shapes(g3d, oracle):
    from generate3d_b200 import raytracing
    rng = np.random.default_rng(11)
    for (H, W, dim) in ((512, 512, 32), (256, 256, 32), (200, 300, 32), (97, 61, 32), (128, 128, 16), (64, 64, 10)):
        poses, depths = [], []
        for q in range(6):
            rot = np.linalg.qr(rng.normal(size=(3, 3)))[0]
            p = np.eye(4, dtype=np.float32)
            p[:3, :3] = rot
            p[:3, 3] = [rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2), -rng.uniform(0.9, 2.0)]
            poses.append(p)
            d = (rng.uniform(size=(H, W)) < 0.002).astype(np.uint8) * 77
            d[H // 3:H // 2, W // 4:W // 2] = 5
            depths.append(d)
        poses, depths = np.stack(poses), np.stack(depths)
        f = 525.0 * W / 512
        r = raytracing.raycast_batch(depths, poses, dim=dim, focal=f, outputs=("occ_bits", "occ_f32", "index_map"))
        for q in range(6):
            want_im = oracle.compute_index_mapping(poses[q], dim, f, W / 2, H / 2, 511)
            want = oracle.raycast_depth(depths[q], poses[q], dim, f, W / 2, H / 2, 511)
            assert np.array_equal(r["index_map"][q].cpu().numpy(), want_im), (H, W, dim, q)
            assert np.array_equal(r["occ_f32"][q].cpu().numpy().astype(np.uint8), want), (H, W, dim, q)
            bits = np.unpackbits(r["occ_bits"][q].cpu().numpy().view(np.uint8), bitorder="little")[:dim ** 3]
            assert np.array_equal(bits.reshape(dim, dim, dim), want)


def test_raycast_edge_cases(g3d, oracle):
    from generate3d_b200 import raytracing
    pose = S.view_pose(2)
    # empty image -> empty grid; full image -> every voxel with a non-empty box
    for fill in (0, 9):
        d = np.full((1, 512, 512), fill, np.uint8)
        r = raytracing.raycast_batch(d, pose[None], outputs=("occ_f32",))
        want = oracle.raycast_depth(d[0], pose, 32, 525.0, 256, 256, 511)
        assert np.array_equal(r["occ_f32"][0].cpu().numpy().astype(np.uint8), want)
    # camera inside the grid (z = 0 crossings -> inf/nan pixels) still matches the x86 conversion rules
    p = np.eye(4, dtype=np.float32)
    d = np.full((1, 512, 512), 3, np.uint8)
    r = raytracing.raycast_batch(d, p[None], outputs=("occ_f32", "index_map"))
    assert np.array_equal(r["index_map"][0].cpu().numpy(), oracle.compute_index_mapping(p, 32, 525.0, 256, 256, 511))
    assert np.array_equal(r["occ_f32"][0].cpu().numpy().astype(np.uint8), oracle.raycast_depth(d[0], p, 32, 525.0, 256, 256, 511))
    # n_view = 0
    r = raytracing.raycast_batch(np.zeros((0, 64, 64), np.uint8), np.zeros((0, 4, 4), np.float32), outputs=("occ_bits",))
    assert tuple(r["occ_bits"].shape) == (0, 1024)


def test_raycast_host_abi(g3d, golden_dir):
    z = np.load(os.path.join(golden_dir, "raycast_256.npz"))
    L = g3d.lib()
    ctx = C.c_void_p()
    g3d.check(L.g3d_context_create(0, C.byref(ctx)))
    depth = np.ascontiguousarray(z["depth"])
    pose = np.ascontiguousarray(z["pose"], np.float32)
    bits = np.empty((20, 1024), np.uint32)
    g3d.check(L.g3d_raycast_batch_host(ctx, depth.ctypes.data, pose.ctypes.data, 20, 256, 256, 262.5, 128.0, 128.0,
                                       511, 32, bits.ctypes.data, None, None))
    L.g3d_context_destroy(ctx)
    assert np.array_equal(bits.view(np.uint8).reshape(20, -1), z["occ_bits"])


def test_dataset_from_tensors_feeds_trainer_contract(g3d):
    """DatasetLoad hands [1,32,32,32] CUDA tensors under the reference's dict keys (dataset_loader.py:173-189)."""
    from generate3d_b200 import dataset_loader, dfgen, raytracing
    V, voff, T, toff = S.table_batch(4, 4)
    r = dfgen.tdf_batch(V, voff, T, toff, outputs=("tdf", "occ_f32"))
    d, p = S.depth_views(4, 256, 256)
    rc = raytracing.raycast_batch(d, p, focal=262.5, outputs=("occ_f32",))
    ds = dataset_loader.DatasetLoad.from_tensors(["m%d" % i for i in range(4)], rc["occ_f32"], r["occ_f32"], r["tdf"])
    assert len(ds) == 4
    item = ds[2]
    assert item["name"] == "m2"
    for k in ("occ_grid", "occ_gt", "df_gt"):
        assert item[k].is_cuda and tuple(item[k].shape) == (1, 32, 32, 32) and item[k].dtype == torch.float32
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=2, num_workers=0)))
    assert tuple(batch["df_gt"].shape) == (2, 1, 32, 32, 32) and batch["df_gt"].is_cuda
    assert float(batch["df_gt"].max()) <= 3.0

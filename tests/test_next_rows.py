"""occupancy -> chamfer DF and the colour raycast, vs the reference's own outputs
(tests/golden/occ_df.npz, raycast_color.npz from make_golden_next.py) and the numpy restatements."""
import json
import os

import numpy as np
import pytest
import torch

from generate3d_b200 import synthetic as S


def test_occ_to_df_oracle_vs_reference(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "occ_df.npz"))
    for q in range(z["occ"].shape[0]):
        want = z["df_false"][q][0]                                            # [z][y][x]
        assert np.array_equal(oracle.occ_to_df(z["occ"][q], 3.0).view(np.uint32), want.view(np.uint32)), q
        occ_pred = (1.0 / (1.0 + np.exp(-z["logits"][q].astype(np.float32))) >= 0.5).astype(np.uint8)
        assert np.array_equal(oracle.occ_to_df(occ_pred, 3.0), z["df_true"][q][0]), q
        xyz = np.transpose(z["occ"][q], (2, 1, 0))
        assert np.array_equal(oracle.occ_to_df(xyz, 3.0, over=4.0), z["df_script"][q]), q


def test_raycast_color_oracle_vs_reference(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "raycast_color.npz"))
    for q in range(z["pose"].shape[0]):
        got = oracle.raycast_color(z["depth"][q], z["color"][q], z["pose"][q])
        assert np.array_equal(got, z["grid_xyz"][q]), q


@pytest.mark.gpu
def test_occ_to_df_gpu_vs_reference(g3d, oracle, golden_dir):
    from generate3d_b200 import data_utils
    z = np.load(os.path.join(golden_dir, "occ_df.npz"))
    occ = torch.from_numpy(z["occ"].astype(np.float32)).cuda().unsqueeze(1)  # [B,1,D,H,W]
    dfs = data_utils.occs_to_dfs(occ, 3.0, pred=False)
    assert tuple(dfs.shape) == (occ.shape[0], 1, 32, 32, 32)
    assert np.array_equal(dfs.cpu().numpy().view(np.uint32), z["df_false"].view(np.uint32))
    one = data_utils.occ_to_df(occ[0], 3.0, pred=False)
    assert np.array_equal(one.cpu().numpy(), z["df_false"][0])
    lg = torch.from_numpy(z["logits"]).cuda().unsqueeze(1)
    assert np.array_equal(data_utils.occs_to_dfs(lg, 3.0, pred=True).cpu().numpy(), z["df_true"])
    for q in (0, 4):
        xyz = torch.from_numpy(np.transpose(z["occ"][q], (2, 1, 0)).astype(np.float32).copy()).cuda()
        assert np.array_equal(data_utils.occ_to_df_xyz(xyz, 3.0).cpu().numpy(), z["df_script"][q])
    # other sizes against the restatement
    rng = np.random.default_rng(5)
    for dim in (8, 20, 40):
        g = (rng.uniform(size=(3, dim, dim, dim)) < 0.01).astype(np.float32)
        got = data_utils.occs_to_dfs(torch.from_numpy(g).cuda().unsqueeze(1), 2.5, pred=False).cpu().numpy()
        for q in range(3):
            assert np.array_equal(got[q, 0], oracle.occ_to_df(g[q], 2.5)), (dim, q)


@pytest.mark.gpu
def test_raycast_color_gpu_vs_reference(g3d, oracle, golden_dir, tmp_path):
    from PIL import Image
    from generate3d_b200 import raytracing
    z = np.load(os.path.join(golden_dir, "raycast_color.npz"))
    g = raytracing.raycast_color_batch(z["depth"], z["color"], z["pose"])
    assert np.array_equal(g.cpu().numpy().astype(np.uint16), z["grid_xyz"])
    Image.fromarray(z["depth"][0], mode="L").save(tmp_path / "d.png")
    Image.fromarray(z["color"][0], mode="RGB").save(tmp_path / "c.png")
    with open(tmp_path / "p.json", "w") as f:
        json.dump({"pose": S.pose_json_payload(0)}, f)
    raytracing.raycast(str(tmp_path / "o.txt"), str(tmp_path / "p.json"), str(tmp_path / "c.png"), str(tmp_path / "d.png"))
    assert (tmp_path / "o.txt").read_bytes() == bytes(z["first_txt"])


def test_raymarch_oracle_vs_reference(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "raymarch.npz"))
    cam = json.loads(str(z["cam"]))["intrinsic"]
    occ, rgb = oracle.raymarch(z["depth"], z["color"], cam["cx"], cam["cy"], cam["fx"])
    assert np.array_equal(occ, z["occ"]) and np.array_equal(rgb, z["rgb"])


@pytest.mark.gpu
def test_raymarch_gpu_vs_reference(g3d, oracle, golden_dir, tmp_path):
    from PIL import Image
    from generate3d_b200 import raytracing_approach1 as ra
    z = np.load(os.path.join(golden_dir, "raymarch.npz"))
    (tmp_path / "cam.json").write_text(str(z["cam"]))
    Image.fromarray(z["depth"], mode="L").save(tmp_path / "d.png")
    Image.fromarray(z["color"], mode="RGB").save(tmp_path / "c.png")
    ra.raycast_to_txt(str(tmp_path / "o.txt"), str(tmp_path / "cam.json"), str(tmp_path / "c.png"), str(tmp_path / "d.png"))
    assert (tmp_path / "o.txt").read_bytes() == bytes(z["txt"])
    cam = ra.load_camera(str(tmp_path / "cam.json"))
    occ, rgb = ra.raymarch(z["depth"], z["color"], cam)
    assert np.array_equal(occ.cpu().numpy(), z["occ"]) and np.array_equal(rgb.cpu().numpy(), z["rgb"])
    # a reference-native 512x512 view against the restatement
    d = S.depth_view(3, 512, 512)
    c = np.random.default_rng(2).integers(0, 256, (512, 512, 3)).astype(np.uint8)
    cam2 = ra.Camera(center=(256, 256), focal=(525.0, 525.0), z_near=0.5, z_far=1.5)
    occ2, rgb2 = ra.raymarch(d, c, cam2)
    wo, wr = oracle.raymarch(d, c, 256, 256, 525.0)
    assert np.array_equal(occ2.cpu().numpy(), wo) and np.array_equal(rgb2.cpu().numpy(), wr)


def test_txt_to_mesh_ply_bytes(golden_dir, tmp_path):
    """voxel_grid.txt_to_mesh: the cube-per-voxel PLY is byte-identical to the reference's."""
    from generate3d_b200 import voxel_grid
    z = np.load(os.path.join(golden_dir, "ply.npz"))
    (tmp_path / "g.txt").write_bytes(bytes(z["txt"]))
    voxel_grid.txt_to_mesh(str(tmp_path / "g.txt"), str(tmp_path / "g.ply"))
    assert (tmp_path / "g.ply").read_bytes() == bytes(z["ply"])


def test_zbuffer_index_map_oracle_vs_reference(oracle, golden_dir):
    """Projection layer's z-buffered index map (Network/perspective_projection.py:65-128): the numpy restatement
    against the reference's own output on 3 grids x 3 poses (tests/golden/make_golden_projection.py)."""
    z = np.load(os.path.join(golden_dir, "projection.npz"))
    for g in range(z["logits"].shape[0]):
        mask = 1.0 / (1.0 + np.exp(-z["logits"][g].astype(np.float64))) > 0.2
        for q in range(z["poses"].shape[1]):
            got = oracle.zbuffer_index_map(mask, z["poses"][g, q])
            assert np.array_equal(got, z["index_map"][g, q].astype(np.int64)), (g, q)
    assert (z["index_map"][2] == -1).all()                                    # empty grid: no pixel is hit


@pytest.mark.gpu
def test_zbuffer_index_map_gpu_vs_reference(g3d, oracle, golden_dir):
    """ProjectionHelper drop-in: batch, n-view and single-view entry points against the reference's maps and
    projected images, plus random occupancy / random poses against the oracle (ties in depth between voxels
    sharing their nearest corner are the rule, not the exception)."""
    from generate3d_b200 import perspective_projection as pp
    z = np.load(os.path.join(golden_dir, "projection.npz"))
    helper = pp.ProjectionHelper()
    logits = torch.from_numpy(z["logits"]).cuda().unsqueeze(1)                # [B,1,D,H,W]
    poses = torch.from_numpy(z["poses"]).cuda()
    maps, projs = helper.project_batch_n_views(logits, poses)
    assert maps.dtype == torch.int64 and tuple(maps.shape) == (3, 3, 512 * 512)
    assert np.array_equal(maps.cpu().numpy(), z["index_map"].astype(np.int64))
    assert np.array_equal(projs.cpu().numpy().view(np.uint32), z["proj"].view(np.uint32))
    one = helper.compute_index_mapping(logits[0], poses[0, 1])
    assert np.array_equal(one.cpu().numpy(), z["index_map"][0, 1])
    nv = helper.index_mapping_n_views(logits[1], poses[1])
    assert np.array_equal(nv.cpu().numpy(), z["index_map"][1])
    # random grids and poses (camera far, near and inside the grid)
    rng = np.random.default_rng(5)
    for trial in range(4):
        occ = rng.uniform(size=(32, 32, 32)) < (0.002, 0.02, 0.2, 0.6)[trial]
        rot = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        p = np.eye(4, dtype=np.float32)
        p[:3, :3] = rot
        p[:3, 3] = [rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2), -(2.0, 1.4, 0.9, 0.3)[trial]]
        lg = torch.from_numpy(np.where(occ, 3.0, -3.0).astype(np.float32)).cuda()
        got = helper.compute_index_mapping(lg, torch.from_numpy(p)).cpu().numpy()
        assert np.array_equal(got, oracle.zbuffer_index_map(occ, p)), trial
    # gradient scatter keeps the reference's shape contract
    g = helper.copy_grad_n_views_occ(maps[0], torch.ones_like(projs[0]))
    assert tuple(g.shape) == (32 ** 3,)
    L = g3d.lib()
    assert L.g3d_zbuffer_index_map(None, None, 1, 1, 512, 512, 525.0, 256.0, 256.0, 511, 32, None, None) == -1


def test_pack_occupancy_mask_bit_order():
    """Host-side packing used by the ProjectionHelper mirror: bit lin & 31 of word lin >> 5, ragged tail padded."""
    from generate3d_b200.perspective_projection import pack_occupancy_mask
    rng = np.random.default_rng(1)
    for n in (32, 64, 100, 32768):
        m = rng.uniform(size=(3, n)) < 0.3
        got = pack_occupancy_mask(torch.from_numpy(m)).numpy().view(np.uint32)
        pad = (-n) % 32
        want = np.packbits(np.pad(m, ((0, 0), (0, pad))).reshape(3, -1, 32), axis=-1, bitorder="little").view(np.uint32).reshape(3, -1)
        assert got.shape == want.shape and np.array_equal(got, want), n

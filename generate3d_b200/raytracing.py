"""Drop-in for src/scripts/raytracing.py: same function names, arguments and
outputs, but the projection and the 32,768-iteration Python loop run as one
CUDA kernel (csrc/g3d_raycast.cu) through the C ABI (g3d_raycast_batch)."""
import json
import os
import pathlib

import numpy as np
import torch

from . import _lib, config
from .dfgen import _device, _to_dev


def make_intrinsic():
    """raytracing.py:27-33."""
    intrinsic = torch.eye(4)
    intrinsic[0][0] = config.focal
    intrinsic[1][1] = config.focal
    intrinsic[0][2] = config.render_img_width / 2
    intrinsic[1][2] = config.render_img_height / 2
    return intrinsic


def make_world_to_grid():
    """raytracing.py:35-40."""
    world_to_grid = torch.eye(4)
    world_to_grid[0][3] = world_to_grid[1][3] = world_to_grid[2][3] = 0.5
    world_to_grid *= config.vox_dim
    world_to_grid[3][3] = 1.0
    return world_to_grid


def load_pose(pose_file):
    """Network/dataset_loader.py:70-81: column-major JSON -> 4x4 f32, camera patched."""
    with open(pose_file) as f:
        data = json.load(f)
    pose = np.transpose(np.reshape(np.asarray(data["pose"], np.float64), (4, 4)))
    pose = torch.from_numpy(pose).float()
    pose[2, 3] = -config.cam_depth
    pose[1, 3] = 0.1
    return pose


def raycast_batch(depth, poses, dim=None, focal=None, cx=None, cy=None, clamp_max=None,
                  outputs=("occ_bits",), device=None):
    """Batch front end: depth u8 [V,H,W], poses f32 [V,4,4] (already load_pose-patched).
    Returns dict with any of occ_bits i32 [V, dim^3/32], occ_f32 f32 [V,dim,dim,dim] ([z][y][x]),
    index_map i64 [V,4,dim^3]. depth may be None when only index_map is requested."""
    device = _device(device if device is not None else (poses.device if torch.is_tensor(poses) and poses.is_cuda else None))
    with torch.cuda.device(device):
        dim = config.vox_dim if dim is None else dim
        poses = _to_dev(poses, torch.float32, device).reshape(-1, 16)
        V = poses.shape[0]
        if depth is not None:
            depth = _to_dev(depth, torch.uint8, device)
            H, W = int(depth.shape[-2]), int(depth.shape[-1])
        else:
            H, W = config.render_img_height, config.render_img_width
        focal = config.focal if focal is None else focal
        cx = W / 2 if cx is None else cx          # make_intrinsic: render_img_width / 2
        cy = H / 2 if cy is None else cy
        clamp_max = config.index_clamp_max if clamp_max is None else clamp_max
        nvox = dim ** 3
        res = {}
        if "occ_bits" in outputs:
            res["occ_bits"] = torch.empty((V, (nvox + 31) // 32), dtype=torch.int32, device=device)
        if "occ_f32" in outputs:
            res["occ_f32"] = torch.empty((V, dim, dim, dim), dtype=torch.float32, device=device)
        if "index_map" in outputs:
            res["index_map"] = torch.empty((V, 4, nvox), dtype=torch.int64, device=device)
        ptr = lambda k: res[k].data_ptr() if k in res else None
        _lib.check(_lib.lib().g3d_raycast_batch(
            depth.data_ptr() if depth is not None else None, poses.data_ptr(), V, H, W, focal, cx, cy,
            clamp_max, dim, ptr("occ_bits"), ptr("occ_f32"), ptr("index_map"),
            torch.cuda.current_stream(device).cuda_stream))
        return res


def compute_index_mapping(world_to_camera, grid_to_world, device):
    """raytracing.py:43-82 -> LongTensor [4 * vox_dim^3] = [umin | vmin | umax | vmax].

    grid_to_world is accepted for signature compatibility; like the reference's only caller
    (raytracing.py:124) it must be inverse(make_world_to_grid())."""
    expect = torch.inverse(make_world_to_grid())
    if grid_to_world is not None and not torch.equal(grid_to_world.detach().cpu().float(), expect):
        raise ValueError("compute_index_mapping: only grid_to_world = inverse(make_world_to_grid()) is supported")
    r = raycast_batch(None, world_to_camera.reshape(1, 4, 4), outputs=("index_map",),
                      device=device if torch.device(device).type == "cuda" else None)
    return torch.flatten(r["index_map"][0])


def occ_rows(occ_zyx, color=(169, 0, 255)):
    """Rows 'x y z r g b' in the reference's order: grid transposed to [x][y][z], np.where order
    (raytracing.py:135-144)."""
    a = occ_zyx.detach().cpu().numpy() if torch.is_tensor(occ_zyx) else np.asarray(occ_zyx)
    x, y, z = np.where(np.transpose(a, (2, 1, 0)) == 1)
    return "".join("%d %d %d %d %d %d\n" % (i, j, k, *color) for i, j, k in zip(x, y, z))


def raycast_color_batch(depth, color, poses, dim=None, focal=None, cx=None, cy=None, clamp_max=None, device=None):
    """Colour raycast for a batch: depth u8 [V,H,W], color u8 [V,H,W,3], poses f32 [V,4,4] ->
    int16 [V, dim, dim, dim, 3] in [x][y][z][rgb] order, 256 = no valid pixel (raytracing.py:85-120)."""
    device = _device(device)
    with torch.cuda.device(device):
        dim = config.vox_dim if dim is None else dim
        poses = _to_dev(poses, torch.float32, device).reshape(-1, 16)
        depth = _to_dev(depth, torch.uint8, device)
        color = _to_dev(color, torch.uint8, device)
        V, H, W = poses.shape[0], int(depth.shape[-2]), int(depth.shape[-1])
        focal = config.focal if focal is None else focal
        cx = W / 2 if cx is None else cx
        cy = H / 2 if cy is None else cy
        clamp_max = config.index_clamp_max if clamp_max is None else clamp_max
        out = torch.empty((V, dim, dim, dim, 3), dtype=torch.int16, device=device)
        _lib.check(_lib.lib().g3d_raycast_color_batch(depth.data_ptr(), color.data_ptr(), poses.data_ptr(), V, H, W,
                                                      focal, cx, cy, clamp_max, dim, out.data_ptr(),
                                                      torch.cuda.current_stream(device).cuda_stream))
        return out


def raycast(txt_file, pose_file, color_file, depth_file):
    """raytracing.py:85-120 (the colour variant raytracing_color.py enables): rows 'x y z r g b' for every voxel
    that saw at least one valid-depth pixel."""
    from PIL import Image
    color = torch.from_numpy(np.array(Image.open(color_file))[..., :3].copy())
    depth = torch.from_numpy(np.array(Image.open(depth_file)))
    pose = load_pose(pose_file)
    g = raycast_color_batch(depth.to(torch.uint8).unsqueeze(0), color.unsqueeze(0), pose.unsqueeze(0),
                            cx=config.render_img_width / 2, cy=config.render_img_height / 2)[0].cpu().numpy()
    pos = np.argwhere(np.all(g < 256, axis=3))
    with open(txt_file, "w") as f:
        for i, j, k in pos:
            f.write("%d %d %d %d %d %d\n" % (i, j, k, g[i, j, k, 0], g[i, j, k, 1], g[i, j, k, 2]))


def raycast_depth(txt_file, pose_file, depth_file):
    """raytracing.py:122-144: depth PNG + pose JSON -> occupancy txt."""
    from PIL import Image
    depth = torch.from_numpy(np.array(Image.open(depth_file)))
    if depth.dim() != 2:
        raise ValueError("depth image must be single-channel ('L')")
    pose = load_pose(pose_file)
    r = raycast_batch(depth.to(torch.uint8).unsqueeze(0), pose.unsqueeze(0), outputs=("occ_f32",),
                      cx=config.render_img_width / 2, cy=config.render_img_height / 2)
    with open(txt_file, "w") as f:
        f.write(occ_rows(r["occ_f32"][0]))


def generate_raycasted_model(synset_id, model_id, renderings_root="/media/sda2/shapenet/shapenet-renderings/",
                             out_root="/media/sda2/shapenet/shapenet-raytraced/", write_ply=True):
    """raytracing.py:147-166: view 00 of a rendered model -> <out_root>/<synset>/<model>.txt (+ the cube PLY)."""
    from .voxel_grid import txt_to_mesh
    outdir = os.path.join(out_root, synset_id)
    pathlib.Path(outdir).mkdir(parents=True, exist_ok=True)
    base = os.path.join(renderings_root, synset_id, model_id)
    voxel_file = os.path.join(outdir, model_id + ".ply")
    voxel_txt_file = os.path.join(outdir, model_id + ".txt")
    if os.path.exists(voxel_txt_file) and (os.path.exists(voxel_file) or not write_ply):
        print(synset_id, " : ", model_id, " already exists. Skipping......")
        return
    raycast_depth(voxel_txt_file, os.path.join(base, "pose", "pose00.json"),
                  os.path.join(base, "depth", "depth00.png"))
    if write_ply:
        txt_to_mesh(voxel_txt_file, voxel_file)

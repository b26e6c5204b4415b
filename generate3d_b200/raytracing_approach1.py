"""Drop-in for the per-pixel ray march of src/scripts/raytracing_approach1.py (+ the bits of voxel_grid.py and
camera.py it needs). SURVEY.md section 8(f) rank 3: kept for completeness, raytracing.py superseded it."""
import json

import numpy as np
import torch

from . import _lib, config
from .dfgen import _device


class Camera:
    """camera.py:13-24,60-85: intrinsics from JSON; z_near / z_far are forced to 0.5 / 1.5 there."""

    def __init__(self, center=(256, 256), focal=(525.0, 525.0), z_near=0.01, z_far=1000.0):
        self.center, self.focal, self.z_near, self.z_far = list(center), list(focal), z_near, z_far

    def load_from_json_file(self, filename):
        with open(filename) as f:
            data = json.load(f)
        self.center = [data["intrinsic"]["cx"], data["intrinsic"]["cy"]]
        self.focal = [data["intrinsic"]["fx"], data["intrinsic"]["fy"]]
        self.z_near, self.z_far = 0.5, 1.5                       # camera.py:70-71
        return self


def load_camera(json_file):
    return Camera().load_from_json_file(json_file)


def raymarch(depth, color, cam, dim=None, device=None):
    """depth u8 [H,W], color u8 [H,W,3] -> (occ u8 [dim,dim,dim], rgb u8 [dim,dim,dim,3]) in [x][y][z] order."""
    device = _device(device)
    dim = config.vox_dim if dim is None else dim
    with torch.cuda.device(device):
        d = torch.as_tensor(np.ascontiguousarray(depth), dtype=torch.uint8).to(device)
        c = torch.as_tensor(np.ascontiguousarray(color)[..., :3].copy(), dtype=torch.uint8).to(device)
        H, W = int(d.shape[0]), int(d.shape[1])
        occ = torch.empty((dim, dim, dim), dtype=torch.uint8, device=device)
        rgb = torch.empty((dim, dim, dim, 3), dtype=torch.uint8, device=device)
        scratch = torch.empty((dim ** 3,), dtype=torch.int32, device=device)
        _lib.check(_lib.lib().g3d_raymarch(d.data_ptr(), c.data_ptr(), H, W, float(cam.center[0]), float(cam.center[1]),
                                           float(cam.focal[0]), float(cam.z_near), float(cam.z_far), dim, occ.data_ptr(),
                                           rgb.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream(device).cuda_stream))
        return occ, rgb


def raycast_to_txt(txt_file, cam_file, color_file, depth_file):
    """raytracing_approach1.py: load_camera + create_voxel_grid + raycast + VoxelGrid.save_vox in one call."""
    from PIL import Image
    cam = load_camera(cam_file)
    occ, rgb = raymarch(np.array(Image.open(depth_file)), np.array(Image.open(color_file)), cam)
    occ, rgb = occ.cpu().numpy(), rgb.cpu().numpy()
    with open(txt_file, "w") as f:
        for i, j, k in zip(*np.where(occ == 1)):                 # voxel_grid.py:111-117
            f.write("%d %d %d %d %d %d\n" % (i, j, k, rgb[i, j, k, 0], rgb[i, j, k, 1], rgb[i, j, k, 2]))

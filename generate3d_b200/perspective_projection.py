"""Drop-in for the data side of src/scripts/Network/perspective_projection.py: `ProjectionHelper` with the
reference's method names and outputs, the z-buffered voxel -> pixel index map computed by one CUDA launch per
batch (csrc/g3d_next.cu, g3d_zbuffer_index_map) instead of a Python loop over the occupied voxels with a slice
assignment per voxel. The autograd wrapper and the PNG/PLY visualisers of the reference are not part of the
data path and are not mirrored."""
import torch

from . import _lib, config
from .dfgen import _device, _to_dev
from .raytracing import make_world_to_grid  # noqa: F401  (same function in both reference modules)


def make_intrinsic():
    """perspective_projection.py:21-27 (image size, not render size)."""
    intrinsic = torch.eye(4)
    intrinsic[0][0] = config.focal
    intrinsic[1][1] = config.focal
    intrinsic[0][2] = config.img_width / 2
    intrinsic[1][2] = config.img_height / 2
    return intrinsic


def pack_occupancy_mask(mask):
    """bool [..., dim^3] (flattened [z][y][x]) -> i32 [..., ceil(dim^3/32)] with bit lin&31 of word lin>>5."""
    n = mask.shape[-1]
    pad = (-n) % 32
    if pad:
        mask = torch.nn.functional.pad(mask, (0, pad))
    w = mask.reshape(*mask.shape[:-1], -1, 32).to(torch.int64)
    weights = (torch.ones(32, dtype=torch.int64, device=mask.device) << torch.arange(32, device=mask.device))
    return (w * weights).sum(-1).to(torch.int32)               # wraps to the signed word with the same bits


class ProjectionHelper:
    """perspective_projection.py:44-160. Tensors live on the CUDA device; there is no CPU path."""

    def __init__(self, device=None):
        self.device = _device(device)
        self.grid_to_world = torch.inverse(make_world_to_grid()).to(self.device)
        self.intrinsic = make_intrinsic().to(self.device)
        self.depth_min = config.znear
        self.depth_max = config.zfar
        self.image_dims = [config.img_width, config.img_height]
        self.volume_dims = [config.vox_dim, config.vox_dim, config.vox_dim]
        self.voxel_size = 1 / config.vox_dim

    def depth_to_skeleton(self, ux, uy, depth):
        x = (ux - self.intrinsic[0][2]) / self.intrinsic[0][0]
        y = (uy - self.intrinsic[1][2]) / self.intrinsic[1][1]
        return torch.Tensor([depth * x, depth * y, depth])

    def skeleton_to_depth(self, p):
        x = (p[0] * self.intrinsic[0][0]) / p[2] + self.intrinsic[0][2]
        y = (p[1] * self.intrinsic[1][1]) / p[2] + self.intrinsic[1][2]
        return torch.Tensor([x, y, p[2]])

    # ---- index maps ------------------------------------------------------------------
    def _occupancy_bits(self, occ):
        """[..., dim^3 logits] -> packed sigmoid(occ) > 0.2 mask (perspective_projection.py:67-69)."""
        occ = _to_dev(occ, torch.float32, self.device)
        return pack_occupancy_mask(torch.gt(torch.sigmoid(occ), 0.2))

    def index_mapping_batch_n_views(self, occ_batch, poses_batch):
        """occ_batch [B, ...dim^3 logits], poses_batch [B, V, 4, 4] -> LongTensor [B, V, H*W]
        (perspective_projection.py:134-140); one launch for the whole batch."""
        dim = self.volume_dims[0]
        B, V = int(poses_batch.shape[0]), int(poses_batch.shape[1])
        W, H = self.image_dims
        with torch.cuda.device(self.device):
            bits = self._occupancy_bits(torch.as_tensor(occ_batch).reshape(B, -1)).contiguous()
            poses = _to_dev(poses_batch, torch.float32, self.device).reshape(B, V, 16).contiguous()
            out = torch.empty((B, V, H * W), dtype=torch.int64, device=self.device)
            if B * V:
                stream = torch.cuda.current_stream(self.device).cuda_stream
                _lib.check(_lib.lib().g3d_zbuffer_index_map(
                    bits.data_ptr(), poses.data_ptr(), B, V, H, W, float(config.focal), float(W / 2), float(H / 2),
                    511, dim, out.data_ptr(), stream))
        return out

    def index_mapping_n_views(self, occ_grid, poses):
        """One grid, V poses -> [V, H*W] (perspective_projection.py:130-132)."""
        poses = torch.as_tensor(poses)
        return self.index_mapping_batch_n_views(torch.as_tensor(occ_grid).reshape(1, -1), poses.reshape(1, -1, 4, 4))[0]

    def compute_index_mapping(self, occ_grid, world_to_camera):
        """One grid, one pose -> [H*W]; -1 where no occupied voxel projects (perspective_projection.py:65-128).
        An empty grid gives all -1 (the reference also prints an error)."""
        return self.index_mapping_n_views(occ_grid, torch.as_tensor(world_to_camera).reshape(1, 4, 4))[0]

    # ---- projections ------------------------------------------------------------------
    def compute_projection(self, occ_grid, index_map):
        """Gather of the occupancy values through an index map, 0 where the map is -1
        (perspective_projection.py:142-153; the map is left unchanged)."""
        occ = _to_dev(torch.as_tensor(occ_grid), torch.float32, self.device).reshape(-1)
        index_map = index_map.to(self.device)
        invalid = torch.eq(index_map, -1)
        proj = torch.index_select(occ, 0, index_map.clamp(min=0))
        proj[invalid] = 0
        return proj

    def project_occ_n_views(self, occ_grid, index_maps):
        return torch.stack([self.compute_projection(occ_grid, m) for m in index_maps])

    def project_batch_n_views(self, occ_batch, poses_batch):
        """(index maps [B,V,H*W], projected images [B,V,H*W]) (perspective_projection.py:159-167)."""
        maps = self.index_mapping_batch_n_views(occ_batch, poses_batch)
        B = int(maps.shape[0])
        projs = torch.stack([self.project_occ_n_views(torch.as_tensor(occ_batch)[i], maps[i]) for i in range(B)])
        return maps, projs

    def copy_grad_to_occ(self, index_map, grad):
        """Scatter of an image gradient back to the grid (perspective_projection.py:169-179)."""
        out = torch.zeros(self.volume_dims[0] * self.volume_dims[1] * self.volume_dims[2], dtype=torch.float,
                          device=self.device)
        valid = torch.gt(index_map, -1)
        out[index_map[valid]] = grad[valid].float()
        return out

    def copy_grad_n_views_occ(self, index_maps, grads):
        return torch.mean(torch.stack([self.copy_grad_to_occ(index_maps[i], grads[i]) for i in range(grads.size(0))]), dim=0)

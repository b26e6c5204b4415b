"""Drop-in for src/scripts/Network/dataset_loader.py on the hot path: load_pose,
load_occ, load_df, load_occ_df and DatasetLoad keep their names, arguments and
return shapes, but grids are GPU-resident torch tensors handed straight to the
trainer (trainer_df.py:112-113 calls .to(device), a no-op for them)."""
import os

import numpy as np
import torch

from . import _lib, config
from .dfgen import _device, read_df
from .raytracing import load_pose  # noqa: F401  (same function, dataset_loader.py:70-81)


def load_occ(txt_file, device=None):
    """dataset_loader.py:102-117: rows 'x y z ...' -> f32 [1,D,H,W] with occ[z][y][x] = 1."""
    device = _device(device)
    rows = np.loadtxt(txt_file, dtype=np.int64, ndmin=2)
    d = config.vox_dim
    occ = torch.zeros((d, d, d), dtype=torch.float32, device=device)
    if rows.size:
        idx = torch.from_numpy(rows[:, :3]).to(device)
        occ[idx[:, 2], idx[:, 1], idx[:, 0]] = 1.0
    return occ.unsqueeze(0)


def load_df(file, trunc_dist, device=None):
    """dataset_loader.py:120-141: .df -> f32 [1,D,H,W], values > trunc_dist clamped (positive side only)."""
    device = _device(device)
    sdf, _, _ = read_df(file)
    src = torch.from_numpy(sdf).to(device)
    out = torch.empty_like(src)
    _lib.check(_lib.lib().g3d_tdf_from_sdf(src.data_ptr(), src.numel(), float(trunc_dist), out.data_ptr(), None,
                                           torch.cuda.current_stream(device).cuda_stream))
    return out.unsqueeze(0)


def load_occ_df(filename, trunc_dist, device=None):
    """dataset_loader.py:144-152: .npy chamfer DF -> clamp, transpose(0,2), add channel."""
    device = _device(device)
    df = torch.from_numpy(np.load(filename)).to(device)
    df = torch.where(df > trunc_dist, torch.full_like(df, trunc_dist), df)
    return torch.transpose(df, 0, 2).unsqueeze(0)


class DatasetLoad(torch.utils.data.Dataset):
    """dataset_loader.py:155-189 with the grids resident in HBM.

    Two ways to fill it:
      * DatasetLoad(data_list, truncation, ...) reads the reference's files once (not per item
        per epoch) using `params` (the dict of parameters.json) for the directories;
      * DatasetLoad.from_tensors(...) wraps grids produced on the GPU by tdf_batch /
        raycast_batch -- no files at all.
    Use with DataLoader(num_workers=0): CUDA tensors must not cross a fork.
    """

    def __init__(self, data_list, truncation, n_max_samples=-1, transform=None, params=None, device=None):
        self.data_list = data_list
        self.n_samples = len(data_list)
        self.transform = transform
        self.truncation = truncation
        if n_max_samples != -1:
            self.n_samples = min(self.n_samples, n_max_samples)
        self._grids = None
        self._params = params
        self._device = device

    @classmethod
    def from_tensors(cls, names, occ_grid, occ_gt, df_gt, occ_df_gt=None, truncation=3.0):
        """occ_grid/occ_gt/df_gt: f32 [N,D,H,W] CUDA tensors ([z][y][x])."""
        self = cls([{"synset_id": "", "model_id": n} for n in names], truncation)
        self._grids = {"occ_grid": occ_grid, "occ_gt": occ_gt, "df_gt": df_gt,
                       "occ_df_gt": occ_df_gt if occ_df_gt is not None else None}
        return self

    def __len__(self):
        return self.n_samples

    def _load_files(self):
        if self._params is None:
            import json
            with open("../parameters.json") as f:                 # dataset_loader.py:177
                self._params = json.load(f)
        p = self._params
        keys = ("occ_grid", "occ_gt", "df_gt", "occ_df_gt")
        cols = {k: [] for k in keys}
        for e in self.data_list[:self.n_samples]:
            s, m = e["synset_id"], e["model_id"]
            cols["occ_grid"].append(load_occ(p["shapenet_raytraced"] + s + "/" + m + ".txt", self._device))
            cols["occ_gt"].append(load_occ(p["shapenet_voxelized"] + s + "/" + m + "__0__.txt", self._device))
            cols["df_gt"].append(load_df(p["shapenet_voxelized"] + s + "/" + m + "__0__.df", self.truncation, self._device))
            f = p["shapenet_voxelized"] + s + "/" + m + "_occ_df.npy"
            cols["occ_df_gt"].append(load_occ_df(f, self.truncation, self._device) if os.path.exists(f) else None)
        self._grids = {k: (torch.cat(v, 0) if all(x is not None for x in v) else None) for k, v in cols.items()}

    def __getitem__(self, index):
        if index >= self.n_samples:
            raise IndexError(index)
        if self._grids is None:
            self._load_files()
        g = self._grids
        item = {"name": self.data_list[index]["model_id"]}
        for k in ("occ_grid", "occ_gt", "occ_df_gt", "df_gt"):
            if g.get(k) is not None:
                item[k] = g[k][index].unsqueeze(0)                # [1,D,H,W]
        return item

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
      * DatasetLoad(data_list, truncation, n_max_samples, transform) -- the reference's constructor. The reference
        re-reads four files (and parameters.json) per item per epoch; here every file is read ONCE, in the
        constructor, in the process that builds the data set (`params`: the dict of parameters.json, default
        "../parameters.json" exactly as dataset_loader.py:177 opens it). An item whose files cannot be read does not
        poison the others: the exception is kept and raised when THAT item is asked for, as the reference would;
      * DatasetLoad.from_tensors(...) wraps grids produced on the GPU by tdf_batch / raycast_batch -- no files.
    Every item carries the reference's five keys ('name', 'occ_grid', 'occ_gt', 'occ_df_gt', 'df_gt').
    The tensors are CUDA tensors, which must not cross a fork: use DataLoader(num_workers=0). The reference
    trainers pass num_workers=2 (trainer_df.py:81); asking for an item from a worker process raises a RuntimeError
    that says so instead of CUDA's "cannot re-initialize CUDA in forked subprocess".
    """

    KEYS = ("occ_grid", "occ_gt", "occ_df_gt", "df_gt")

    def __init__(self, data_list, truncation, n_max_samples=-1, transform=None, params=None, device=None, _load=True):
        self.data_list = data_list
        self.n_samples = len(data_list)
        self.transform = transform
        self.truncation = truncation
        if n_max_samples != -1:
            self.n_samples = min(self.n_samples, n_max_samples)
        self._grids = None
        self._errors = {}
        self._params = params
        self._device = device
        if _load:
            self._load_files()

    @classmethod
    def from_tensors(cls, names, occ_grid, occ_gt, df_gt, occ_df_gt=None, truncation=3.0):
        """occ_grid / occ_gt / df_gt: f32 [N,D,H,W] CUDA tensors ([z][y][x]). Without occ_df_gt the chamfer field of
        occ_gt is computed on the GPU -- what occ_to_df.py:16-62 saves as *_occ_df.npy and load_occ_df
        (dataset_loader.py:144-152) clamps at the truncation."""
        self = cls([{"synset_id": "", "model_id": n} for n in names], truncation, _load=False)
        if occ_df_gt is None:
            from .data_utils import _batch
            occ_df_gt = _batch(occ_gt.reshape(-1, *occ_gt.shape[-3:]), truncation, False, truncation)
        self._grids = {"occ_grid": occ_grid, "occ_gt": occ_gt, "df_gt": df_gt, "occ_df_gt": occ_df_gt}
        for k, g in self._grids.items():
            if g.shape[0] != len(names):
                raise ValueError("%s holds %d grids for %d names" % (k, g.shape[0], len(names)))
        return self

    def __len__(self):
        return self.n_samples

    def _load_files(self):
        if self._params is None:
            import json
            with open("../parameters.json") as f:                 # dataset_loader.py:177
                self._params = json.load(f)
        p = self._params
        device = _device(self._device)
        d = config.vox_dim
        self._grids = {k: torch.zeros((self.n_samples, d, d, d), dtype=torch.float32, device=device) for k in self.KEYS}
        for i, e in enumerate(self.data_list[:self.n_samples]):
            s, m = e["synset_id"], e["model_id"]
            vox = p["shapenet_voxelized"] + s + "/" + m
            try:                                                  # the four loads of dataset_loader.py:179-187
                self._grids["occ_grid"][i] = load_occ(p["shapenet_raytraced"] + s + "/" + m + ".txt", device)[0]
                self._grids["occ_gt"][i] = load_occ(vox + "__0__.txt", device)[0]
                self._grids["df_gt"][i] = load_df(vox + "__0__.df", self.truncation, device)[0]
                self._grids["occ_df_gt"][i] = load_occ_df(vox + "_occ_df.npy", self.truncation, device)[0]
            except Exception as err:                              # noqa: BLE001 -- raised again from __getitem__(i)
                self._errors[i] = err

    def __getitem__(self, index):
        if index >= self.n_samples:
            raise IndexError(index)
        if torch.utils.data.get_worker_info() is not None:
            raise RuntimeError("generate3d_b200.DatasetLoad keeps its grids in GPU memory and cannot be read from a "
                               "DataLoader worker process: build the DataLoader with num_workers=0")
        if index in self._errors:
            raise self._errors[index]
        item = {"name": self.data_list[index]["model_id"]}
        for k in self.KEYS:
            item[k] = self._grids[k][index].unsqueeze(0)          # [1,D,H,W]
        return item

"""Drop-in for the hot part of src/scripts/Network/data_utils.py:
preprocess_occ / occ_to_df / occs_to_dfs, and the occ_to_df of src/scripts/occ_to_df.py, run as one CUDA
kernel over the whole batch (g3d_occ_to_df_batch) instead of a Python loop of fancy-indexing ops."""
import torch

from . import _lib


def _batch(occ, trunc, pred, over):
    if not occ.is_cuda:
        raise RuntimeError("generate3d_b200 has no CPU path: occ_to_df needs a CUDA tensor")
    x = occ.contiguous().float()
    n, d = x.shape[0], x.shape[-1]
    out = torch.empty_like(x)
    _lib.check(_lib.lib().g3d_occ_to_df_batch(x.data_ptr(), n, d, 1 if pred else 0, float(trunc), float(over),
                                              out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream))
    return out


def preprocess_occ(occ, pred=True):
    """data_utils.py:100-112: binarise a predicted occupancy ([1,D,H,W]) and return it as [W,H,D]."""
    occ = occ.clone()
    if pred:
        occ = torch.sigmoid(occ)
        m = occ >= 0.5
        occ[m] = 1.0
        occ[~m] = 0.0
    return torch.transpose(occ[0], 0, 2)


def occ_to_df(occ, trunc, pred=True):
    """data_utils.py:115-168: occ [1,D,H,W] -> chamfer DF [1,D,H,W], values above trunc set to trunc."""
    return _batch(occ.reshape(1, *occ.shape[-3:]), trunc, pred, trunc)


def occs_to_dfs(occs, trunc, pred=True):
    """data_utils.py:171-177: batch [B,1,D,H,W] -> [B,1,D,H,W]; one launch instead of B Python calls."""
    b = occs.shape[0]
    return _batch(occs.reshape(b, *occs.shape[-3:]), trunc, pred, trunc).reshape(b, 1, *occs.shape[-3:])


def occ_to_df_xyz(occ_xyz, trunc):
    """src/scripts/occ_to_df.py:16-62: occ [x][y][z] of 0/1 -> DF [x][y][z], values above trunc set to trunc + 1
    (the array that script saves as *_occ_df.npy)."""
    zyx = torch.transpose(occ_xyz, 0, 2).unsqueeze(0)
    return torch.transpose(_batch(zyx, trunc, False, trunc + 1)[0], 0, 2)

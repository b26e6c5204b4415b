"""tdflog definition, mirroring src/scripts/Network/losses.py:7-11, computed by
the fused CUDA epilogue (g3d_tdf_from_sdf) instead of four torch ops."""
import torch

from . import _lib


def apply_log_transform(df):
    """sign(df) * log(|df| + 1) for a CUDA float tensor (losses.py:7-11)."""
    if not df.is_cuda:
        raise RuntimeError("generate3d_b200 has no CPU path: apply_log_transform needs a CUDA tensor")
    src = df.contiguous().float()
    out = torch.empty_like(src)
    L = _lib.lib()
    _lib.check(L.g3d_tdf_from_sdf(src.data_ptr(), src.numel(), float("inf"), None, out.data_ptr(),
                                  torch.cuda.current_stream(src.device).cuda_stream))
    return out

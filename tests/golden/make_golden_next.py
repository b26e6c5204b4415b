"""Golden fixtures for the SURVEY.md section 8(f) "next" rows, produced by the REFERENCE's own code
(build container only, needs /root/reference):

  occ_df.npz       Network/data_utils.py occ_to_df (pred=False and pred=True) and occ_to_df.py occ_to_df
  raycast_color.npz  raytracing.py raycast (colour variant) on two synthetic views
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from make_golden import import_reference, REF  # noqa: E402


class legacy_integer_division:
    """The reference pins torch==1.3.0 (requirements.txt:70), where `/` on integer tensors is integer
    division; data_utils.occ_to_df relies on it (lin_ind / 9, lin_ind_volume / (34*34), ...). Under the
    torch installed here `/` is true division and the function raises. This context manager restores the
    pinned semantics for integer-by-integer `/` only, leaving the reference's code untouched."""

    def __enter__(self):
        import torch
        self.torch = torch
        self.orig = torch.Tensor.__truediv__

        def div(a, b):
            b_int = isinstance(b, int) or (torch.is_tensor(b) and not b.is_floating_point())
            if torch.is_tensor(a) and not a.is_floating_point() and b_int:
                return torch.div(a, b, rounding_mode="trunc")
            return self.orig(a, b)
        torch.Tensor.__truediv__ = div
        return self

    def __exit__(self, *exc):
        self.torch.Tensor.__truediv__ = self.orig


def main():
    import torch
    from generate3d_b200 import synthetic as S
    from oracle import oracle as O
    config, raytracing, dataset_loader, losses = import_reference()
    cwd = os.getcwd()
    os.chdir(REF)
    import data_utils
    import occ_to_df as occ_to_df_script
    os.chdir(cwd)

    # ---- occupancy -> chamfer DF
    grids = []
    v, t = S.table_mesh(6, seed=2)
    grids.append(O.occupancy(O.dfgen(t, v, 32)["sdf"]))                       # GT occupancy of a table, [z][y][x]
    d, p = S.depth_views(1, 256, 256)
    grids.append(O.raycast_depth(d[0], p[0], 32, 262.5))                      # a raycast occupancy
    g = np.zeros((32, 32, 32), np.uint8); grids.append(g.copy())              # empty
    g[:] = 1; grids.append(g.copy())                                          # full
    g[:] = 0; g[0, 0, 0] = 1; g[31, 31, 31] = 1; g[5, 17, 9] = 1; grids.append(g.copy())   # corners + one voxel
    rng = np.random.default_rng(3)
    grids.append((rng.uniform(size=(32, 32, 32)) < 0.002).astype(np.uint8))   # sparse noise
    occ = np.stack(grids)
    out_false, out_true, out_script = [], [], []
    logits = rng.normal(size=occ.shape).astype(np.float32) * 3
    ctx = legacy_integer_division()
    ctx.__enter__()
    for q in range(occ.shape[0]):
        o = torch.from_numpy(occ[q].astype(np.float32)).unsqueeze(0)          # [1, D, H, W]
        out_false.append(data_utils.occ_to_df(o.clone(), 3.0, pred=False).numpy())
        out_true.append(data_utils.occ_to_df(torch.from_numpy(logits[q]).unsqueeze(0).clone(), 3.0, pred=True).numpy())
        xyz = torch.transpose(o[0], 0, 2)                                     # occ_to_df.py is fed [x][y][z] (its __main__)
        out_script.append(occ_to_df_script.occ_to_df(xyz.clone(), 3.0).numpy())
    ctx.__exit__()
    np.savez_compressed(os.path.join(HERE, "occ_df.npz"), occ=occ, logits=logits, df_false=np.stack(out_false),
                        df_true=np.stack(out_true), df_script=np.stack(out_script))
    print("occ_df: unique values", np.unique(np.stack(out_false))[:12])

    # ---- colour raycast
    from PIL import Image
    tmp = tempfile.mkdtemp()
    depths, colors, poses, grids_c = [], [], [], []
    for k in (0, 7):
        dep = S.depth_view(k, 512, 512)
        col = rng.integers(0, 256, (512, 512, 3)).astype(np.uint8)
        col[::7, ::5] = 255
        Image.fromarray(dep, mode="L").save(os.path.join(tmp, "d.png"))
        Image.fromarray(col, mode="RGB").save(os.path.join(tmp, "c.png"))
        with open(os.path.join(tmp, "p.json"), "w") as f:
            json.dump({"pose": S.pose_json_payload(k)}, f)
        raytracing.raycast(os.path.join(tmp, "o.txt"), os.path.join(tmp, "p.json"), os.path.join(tmp, "c.png"),
                           os.path.join(tmp, "d.png"))
        with open(os.path.join(tmp, "o.txt")) as f:
            text = f.read()
        grid = np.full((32, 32, 32, 3), 256, np.uint16)                       # [x][y][z][rgb] as the txt rows index it
        for line in text.splitlines():
            x, y, z, r, g_, b = (int(s) for s in line.split())
            grid[x, y, z] = (r, g_, b)
        depths.append(dep); colors.append(col); poses.append(S.view_pose(k)); grids_c.append(grid)
        if k == 0:
            first_txt = text
    np.savez_compressed(os.path.join(HERE, "raycast_color.npz"), depth=np.stack(depths), color=np.stack(colors),
                        pose=np.stack(poses), grid_xyz=np.stack(grids_c), first_txt=np.frombuffer(first_txt.encode(), np.uint8))
    print("raycast_color: coloured voxels", [(g[..., 0] < 256).sum() for g in grids_c])


if __name__ == "__main__":
    main()

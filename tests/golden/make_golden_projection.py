"""Golden fixtures for the z-buffered voxel -> pixel index map of the reference's projection layer
(src/scripts/Network/perspective_projection.py:65-128), produced by the REFERENCE's own code on CPU torch
(build container only, needs /root/reference):

  projection.npz   occupancy logits [G,32,32,32], poses [G,V,4,4], index maps [G,V,512*512] (int32; -1 = no voxel),
                   projected images [G,V,512*512]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from make_golden import import_reference, REF  # noqa: E402
from make_golden_next import legacy_integer_division  # noqa: E402


def main():
    import torch
    from generate3d_b200 import synthetic as S
    from oracle import oracle as O
    import_reference()
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "Network"))
    import perspective_projection as pp
    os.chdir(cwd)
    helper = pp.ProjectionHelper()

    rng = np.random.default_rng(7)
    grids = []
    v, t = S.table_mesh(6, seed=2)
    occ = O.occupancy(O.dfgen(t, v, 32)["sdf"]).astype(np.float32)             # GT occupancy of a table, [z][y][x]
    grids.append(np.where(occ > 0, 4.0, -4.0).astype(np.float32) + rng.normal(size=occ.shape).astype(np.float32) * 0.3)
    g = np.full((32, 32, 32), -5.0, np.float32)                                # a few isolated voxels incl. corners
    for (z, y, x) in ((0, 0, 0), (31, 31, 31), (5, 17, 9), (5, 17, 10), (16, 16, 16), (16, 16, 17)):
        g[z, y, x] = 3.0
    grids.append(g)
    grids.append(np.full((32, 32, 32), -6.0, np.float32))                      # empty: every pixel -1
    logits = np.stack(grids)
    poses = np.stack([np.stack([S.view_pose(k) for k in (0, 3, 7)]) for _ in range(len(grids))]).astype(np.float32)
    maps, imgs = [], []
    with legacy_integer_division():
        for gi in range(logits.shape[0]):
            o = torch.from_numpy(logits[gi]).unsqueeze(0)                     # [1, D, H, W] as the trainers hold it
            m = helper.index_mapping_n_views(o, torch.from_numpy(poses[gi]))
            p = helper.project_occ_n_views(o, m.clone())
            maps.append(m.numpy().astype(np.int32))
            imgs.append(p.numpy().astype(np.float32))
            print("grid", gi, "pixels hit per view", [(int((mm >= 0).sum())) for mm in maps[-1]])
    np.savez_compressed(os.path.join(HERE, "projection.npz"), logits=logits, poses=poses, index_map=np.stack(maps),
                        proj=np.stack(imgs))


if __name__ == "__main__":
    main()

"""Small end-to-end cases for compute-sanitizer (race / memory checking of the CUDA kernels).

    compute-sanitizer --tool racecheck python tools/sanitize_case.py
    compute-sanitizer --tool memcheck  python tools/sanitize_case.py

Every kernel of the path runs at least once on inputs small enough for the tools' 10-100x slowdown: the tiled and the
scatter band init, 128-thread sweeps (two rounds per queue), 1,024-thread sweeps and the thread-block-cluster sweep,
the exact-mode kernel (TMA bulk copies + mbarrier), finalize, raycast, and the "next" rows. Results are still checked
against the oracle, so a run that the tool slowed into a different interleaving must give the same bits.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from generate3d_b200 import data_utils, dfgen, raytracing, synthetic as S  # noqa: E402
from oracle import oracle as O  # noqa: E402


def check(tag, V, voff, T, toff, dim, mode="faithful"):
    r = dfgen.tdf_batch(V, voff, T, toff, dim=dim, mode=mode, outputs=("sdf", "closest_tri", "occ_bits", "status"))
    torch.cuda.synchronize()
    sdf = r["sdf"].cpu().numpy()
    if mode == "faithful":
        for m in range(len(voff) - 1):
            want = O.dfgen(T[toff[m]:toff[m + 1]], V[voff[m]:voff[m + 1]], dim)["sdf"]
            assert np.array_equal(sdf[m].view(np.uint32), want.view(np.uint32)), (tag, m)
    print("ok", tag, flush=True)


def main():
    small = S.table_batch(6, 2, first_seed=0)
    # tiled band init + 128-thread sweeps forced (the 1,024-mesh batch path)
    os.environ["G3D_INIT_TILED"], os.environ["G3D_SWEEP_THREADS"] = "1", "128"
    check("tiled init + sweep<128>", *small, 32)
    os.environ["G3D_SWEEP_THREADS"] = "256"
    check("tiled init + sweep<256>", *small, 32)
    # scatter init + 1,024-thread sweep (single-mesh path), then the cluster sweep
    os.environ["G3D_INIT_TILED"] = "0"
    os.environ.pop("G3D_SWEEP_THREADS")
    one = S.table_batch(1, 3, first_seed=7)
    check("scatter init + sweep<1024>", *one, 32)
    os.environ["G3D_SWEEP_CLUSTER"] = "2"
    check("cluster sweep x2", *one, 40)
    os.environ["G3D_SWEEP_CLUSTER"] = "4"
    check("cluster sweep x4", *one, 40)
    os.environ.pop("G3D_SWEEP_CLUSTER")
    os.environ.pop("G3D_INIT_TILED")
    check("exact mode", *small, 32, mode="exact")
    d, p = S.depth_views(2, 256, 256)
    rc = raytracing.raycast_batch(d, p, focal=262.5, outputs=("occ_f32", "index_map"))
    torch.cuda.synchronize()
    for q in range(2):
        assert np.array_equal(rc["occ_f32"][q].cpu().numpy().astype(np.uint8), O.raycast_depth(d[q], p[q], 32, 262.5))
    print("ok raycast", flush=True)
    occ = rc["occ_f32"].unsqueeze(1)
    data_utils.occs_to_dfs(occ, 3.0, pred=False)
    col = torch.zeros((2, 256, 256, 3), dtype=torch.uint8, device="cuda")
    raytracing.raycast_color_batch(torch.from_numpy(d).cuda(), col, torch.from_numpy(p).cuda(), focal=262.5)
    torch.cuda.synchronize()
    print("ok next rows", flush=True)


if __name__ == "__main__":
    main()

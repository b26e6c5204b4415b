"""Small driver for ncu captures: runs the hot path a few times on synthetic input, nothing else.

    python tools/profile_case.py tdf  [--meshes 296] [--n 18] [--mode faithful|exact] [--reps 3]
    python tools/profile_case.py raycast [--models 64] [--size 256] [--reps 3]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["tdf", "raycast"])
    ap.add_argument("--meshes", type=int, default=296)
    ap.add_argument("--n", type=int, default=18)
    ap.add_argument("--dim", type=int, default=32)
    ap.add_argument("--mode", default="faithful")
    ap.add_argument("--models", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    from generate3d_b200 import dfgen, raytracing, synthetic as S
    if a.what == "tdf":
        base = min(a.meshes, 37)
        V, voff, T, toff = S.table_batch(base, a.n)
        reps = (a.meshes + base - 1) // base
        Vs, Ts, vo, to = [], [], [0], [0]
        for r in range(reps):
            for m in range(base):
                if len(vo) - 1 >= a.meshes:
                    break
                Vs.append(V[voff[m]:voff[m + 1]]); Ts.append(T[toff[m]:toff[m + 1]])
                vo.append(vo[-1] + len(Vs[-1])); to.append(to[-1] + len(Ts[-1]))
        V, T = np.concatenate(Vs), np.concatenate(Ts)
        dV, dT = torch.from_numpy(V).cuda(), torch.from_numpy(T.view(np.int32)).cuda()
        dvo, dto = torch.tensor(vo).cuda(), torch.tensor(to).cuda()
        out = None
        for _ in range(a.reps):
            out = dfgen.tdf_batch(dV, dvo, dT, dto, dim=a.dim, mode=a.mode, outputs=("sdf", "tdf", "tdflog", "occ_bits"), out=out)
        torch.cuda.synchronize()
        print("tdf ok", a.meshes, "meshes", float(out["sdf"].abs().mean()))
    else:
        d, p = S.depth_views(20, a.size, a.size)
        dd = torch.from_numpy(d).cuda().repeat(a.models, 1, 1)
        dp = torch.from_numpy(p).cuda().repeat(a.models, 1, 1)
        for _ in range(a.reps):
            r = raytracing.raycast_batch(dd, dp, focal=525.0 * a.size / 512, outputs=("occ_bits",))
        torch.cuda.synchronize()
        print("raycast ok", dd.shape[0], "views", int(r["occ_bits"].ne(0).sum()))


if __name__ == "__main__":
    main()

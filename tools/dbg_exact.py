import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
from generate3d_b200 import dfgen, _lib, synthetic as S
def run(meshes, trunc, dim=16):
    p, _ = dfgen.dfgen_geometry(dim)
    p.trunc, p.occ_thresh, p.mode = trunc, 1.0, _lib.MODE_EXACT
    voff, toff = [0], [0]
    for v, t in meshes:
        voff.append(voff[-1] + len(v)); toff.append(toff[-1] + len(t))
    V = np.concatenate([m[0] for m in meshes]); T = np.concatenate([m[1] for m in meshes])
    try:
        r = dfgen.tdf_batch_params(V, np.asarray(voff), T, np.asarray(toff), p, ("sdf", "closest_tri", "status"))
        torch.cuda.synchronize()
        return "ok " + str(r["status"].cpu().tolist())
    except Exception as e:
        return "FAIL " + str(e)[:80]
case = sys.argv[1]
m1, m2 = S.table_mesh(3, 50), S.table_mesh(2, 51, scale=0.7, rot_y=1.0)
bv, bt = S.table_mesh(2, 52); bt = bt.copy(); bt[1, 2] = 99999
cases = {"a": ([m1], 3.0), "b": ([m1], float("inf")), "c": ([m1, m2], 3.0), "d": ([m1, m2], float("inf")), "e": ([(bv, bt)], 3.0), "f": ([m1, m2, (bv, bt)], float("inf")), "g": ([m2], float("inf"))}
print(case, os.environ.get("G3D_EXACT_STRICT"), run(*cases[case]))

"""BASELINE configs[3]: 128^3 TDF of a ~200k-triangle mesh (n=58 -> 201,840 tris). Parity vs the CPU oracle + timing."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generate3d_b200 import dfgen, synthetic as S
from oracle import oracle as O

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else 58
v, t = S.table_mesh(n, seed=0)
print("tris", len(t), "dim", dim)
voff, toff = np.array([0, len(v)]), np.array([0, len(t)])
for mode in ("faithful", "exact"):
    r = dfgen.tdf_batch(v, voff, t, toff, dim=dim, mode=mode, outputs=("sdf", "occ_f32"))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        r = dfgen.tdf_batch(v, voff, t, toff, dim=dim, mode=mode, outputs=("sdf", "occ_f32"), out=r)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print(mode, "%.2f ms per mesh -> %.3g voxels/s" % (ms, dim ** 3 / ms * 1e3))
    if mode == "faithful":
        f_sdf = r["sdf"][0].cpu().numpy(); f_occ = r["occ_f32"][0].cpu().numpy()
    else:
        e_sdf = r["sdf"][0].cpu().numpy(); e_occ = r["occ_f32"][0].cpu().numpy()
t0 = time.perf_counter()
want = O.dfgen(t, v, dim, use_ref=O.ref_available())["sdf"]
print("cpu reference %.2f s" % (time.perf_counter() - t0))
print("faithful bit-equal:", np.array_equal(f_sdf.view(np.uint32), want.view(np.uint32)))
print("exact occupancy == reference occupancy:", np.array_equal(e_occ, f_occ))
band = np.abs(want) <= 3
print("exact vs reference in |d|<=3: mismatches > 1e-5: %d of %d, max %.4g" % ((np.abs(np.abs(e_sdf) - np.abs(want))[band] > 1e-5).sum(), band.sum(), np.abs(np.abs(e_sdf) - np.abs(want))[band].max()))

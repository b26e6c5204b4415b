"""Timing of exact mode on the configs[2] batch (tuning aid): python tools/exact_time.py [meshes]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generate3d_b200 import _lib, dfgen, synthetic as S
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
V, voff, T, toff = S.table_batch(M, 18, 0)
dV, dT = torch.from_numpy(V).cuda(), torch.from_numpy(T.view(np.int32)).cuda()
dvo, dto = torch.from_numpy(voff).cuda(), torch.from_numpy(toff).cuda()
L = _lib.lib()
outs = ("sdf", "tdf", "tdflog", "occ_bits", "status")
o = dfgen.tdf_batch(dV, dvo, dT, dto, mode="exact", outputs=outs)
_lib.check(L.g3d_profile_enable(1))
sm = (C.c_float * 8)()
for rep in range(4):
    dfgen.tdf_batch(dV, dvo, dT, dto, mode="exact", outputs=outs, out=o)
    _lib.check(L.g3d_profile_read(sm, None))
st = np.frombuffer(sm, np.float32)
print("G3D_EXACT_GROUP=%s G3D_EXACT_OCC=%s  prep %.2f index %.2f exact %.2f finalize %.2f ms" % (os.environ.get("G3D_EXACT_GROUP"), os.environ.get("G3D_EXACT_OCC"), st[1], st[3], st[6], st[5]), flush=True)

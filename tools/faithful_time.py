"""Step time, stage times and evaluation counts of faithful mode on the configs[2] batch (tuning aid):
    python tools/faithful_time.py [meshes]"""
import ctypes as C
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generate3d_b200 import dfgen, _lib, synthetic as S
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
V, voff, T, toff = S.table_batch(M, 18, 0)
dV, dT = torch.from_numpy(V).cuda(), torch.from_numpy(T.view(np.int32)).cuda()
dvo, dto = torch.from_numpy(voff).cuda(), torch.from_numpy(toff).cuda()
outs = ("sdf", "tdf", "tdflog", "occ_bits", "status")
L = _lib.lib()
ev = (C.c_uint64 * 4)()
sm = (C.c_float * 8)()
_lib.check(L.g3d_profile_enable(2))
o = dfgen.tdf_batch(dV, dvo, dT, dto, outputs=outs)
torch.cuda.synchronize()
_lib.check(L.g3d_profile_read(None, ev))
_lib.check(L.g3d_profile_enable(1))
for _ in range(3):
    dfgen.tdf_batch(dV, dvo, dT, dto, outputs=outs, out=o)
torch.cuda.synchronize()
_lib.check(L.g3d_profile_read(sm, None))
stages = [round(float(x), 3) for x in sm]
_lib.check(L.g3d_profile_enable(0))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    dfgen.tdf_batch(dV, dvo, dT, dto, outputs=outs, out=o)
b.record()
torch.cuda.synchronize()
print("faithful, %d meshes: %.2f ms/step  evals init %.4g sweep %.4g  stages %s" % (M, a.elapsed_time(b) / 10, ev[0], ev[1], stages), flush=True)

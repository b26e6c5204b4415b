"""Development tool: what the 16 sweeps of the faithful path do on the bench's synthetic meshes, sweep by sweep (CPU replay
with the sweep kernels' bookkeeping; tools/sweep_stats.c).   python tools/sweep_stats.py [n_mesh]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from generate3d_b200 import synthetic as S          # noqa: E402
from oracle import oracle as O                      # noqa: E402

so = "/tmp/sweep_stats.so"
subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                       os.path.join(ROOT, "tools", "sweep_stats.c"), "-lm"])
L = C.CDLL(so)
n_mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = O.dfgen_geometry(32)
tot = np.zeros((16, 8), np.int64)
for m in range(n_mesh):
    v, vo, t, to = S.table_batch(1, 18, first_seed=m)
    out = np.zeros((16, 8), np.int64)
    L.sweep_stats(t.ctypes.data_as(C.c_void_p), C.c_int64(len(t)), v.ctypes.data_as(C.c_void_p),
                  g["origin"].ctypes.data_as(C.c_void_p), C.c_float(float(g["dx"])), 32, 32, 32,
                  out.ctypes.data_as(C.c_void_p))
    tot += out
print("per mesh, averaged over %d meshes of 19,440 triangles at 32^3 (the reference evaluates 7 x 29,791 pairs per sweep)" % n_mesh)
print("  s     visits    w/cand. evals(1-2) seen-before in-band-box evals(1-3) evals(1-4)    changed")
for s in range(16):
    print("%3d" % s, " ".join("%10.0f" % (x / n_mesh) for x in tot[s]))
print("sum", " ".join("%10.0f" % (x / n_mesh) for x in tot.sum(0)))

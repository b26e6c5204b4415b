"""Development tool: per-sweep statistics of the faithful sweeps on the bench's synthetic meshes (CPU replay).
    python tools/sweep_stats.py [n_mesh]"""
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
etot = np.zeros((16, 4), np.int64)
e2tot = np.zeros((16, 8), np.int64)
e3tot = np.zeros((16, 4), np.int64)
for m in range(n_mesh):
    v, vo, t, to = S.table_batch(1, 18, first_seed=m)
    out = np.zeros((16, 8), np.int64)
    ex = np.zeros((16, 4), np.int64)
    e2 = np.zeros((16, 8), np.int64)
    e3 = np.zeros((16, 4), np.int64)
    L.sweep_stats(t.ctypes.data_as(C.c_void_p), C.c_int64(len(t)), v.ctypes.data_as(C.c_void_p),
                  g["origin"].ctypes.data_as(C.c_void_p), C.c_float(float(g["dx"])), 32, 32, 32,
                  out.ctypes.data_as(C.c_void_p), ex.ctypes.data_as(C.c_void_p), e2.ctypes.data_as(C.c_void_p), e3.ctypes.data_as(C.c_void_p))
    tot += out
    e2tot += e2
    e3tot += e3
    etot += ex
print("per mesh averages over", n_mesh)
print("  s  visits  stampact  evalnodes   evals  changes  rounds  actrounds  actlevels")
for s in range(16):
    print("%3d" % s, " ".join("%9.0f" % (x / n_mesh) for x in tot[s]))
print("sum", " ".join("%9.0f" % (x / n_mesh) for x in tot.sum(0)))
print('in-band-box / seen-before / either / dup-of-examined-upwind:')
for s in range(16):
    print('%3d' % s, ' '.join('%9.0f' % (x / n_mesh) for x in etot[s]))
print('sum', ' '.join('%9.0f' % (x / n_mesh) for x in etot.sum(0)))
print('fifo2 / fifo4 / fifo8 / tri-ineq / ti|fifo4 / ti|seen|inb / ti|fifo2:')
for s in range(16):
    print('%3d' % s, ' '.join('%9.0f' % (x / n_mesh) for x in e2tot[s]))
print('sum', ' '.join('%9.0f' % (x / n_mesh) for x in e2tot.sum(0)))
print('after the dup-of-examined rule: evals left / in band box / 1-bit interior flag / 6-bit flags')
for s in range(16):
    print('%3d' % s, '%9.0f' % (e2tot[s][7] / n_mesh), ' '.join('%9.0f' % (x / n_mesh) for x in e3tot[s][:3]))
print('sum', '%9.0f' % (e2tot[:, 7].sum() / n_mesh), ' '.join('%9.0f' % (x / n_mesh) for x in e3tot.sum(0)[:3]))

"""CPU check of the exactness argument behind the sweep kernels' shortcuts (no GPU): tests/sweep_rules_model.c replays
the 16 sweeps with the kernels' rules -- change codes and re-examination thresholds, candidates equal to a "met"
neighbour's triangle, band-box flags, speculative sweeps with a level-by-level pass over the would-change nodes only --
and must reproduce the oracle's phi and closest_tri bit for bit while making a fraction of its evaluations."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("sweep_rules") / "sweep_rules_model.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "sweep_rules_model.c"), "-lm"])
    lib = C.CDLL(so)
    lib.sweep_rules_model.restype = C.c_int64
    return lib


def _cases():
    import sys
    sys.path.insert(0, ROOT)
    from generate3d_b200 import synthetic as S
    rng = np.random.default_rng(7)
    v, t = S.table_mesh(4, seed=3)
    yield "table 960 tris, 24^3", t, v, np.float32([-0.5, -0.5, -0.5]), np.float32(1.0 / 24), (24, 24, 24)
    v, t = S.table_mesh(9, seed=1, scale=0.8, rot_y=0.7)
    yield "table 4,860 tris, 32^3", t, v, np.float32([-0.5, -0.5, -0.5]), np.float32(1.0 / 32), (32, 32, 32)
    for q in range(6):
        nv, nt = int(rng.integers(3, 200)), int(rng.integers(1, 400))
        v = rng.uniform(-0.7, 0.7, (nv, 3)).astype(np.float32)
        t = rng.integers(0, nv, (nt, 3)).astype(np.uint32)
        if q % 3 == 2:                                   # large triangles crossing the grid, repeated vertices
            v = rng.uniform(-3, 3, (12, 3)).astype(np.float32)
            t = rng.integers(0, 12, (7, 3)).astype(np.uint32)
            t[0, 1] = t[0, 0]
        dims = tuple(int(x) for x in rng.integers(2, 23, 3))
        dx = np.float32(rng.uniform(0.5, 1.5) / max(dims))
        origin = np.float32(-0.5) + rng.uniform(-0.1, 0.1, 3).astype(np.float32)
        yield "soup %d %s" % (q, dims), t, v, origin, dx, dims


@pytest.mark.parametrize("spec_from", [99, 8, 1])
def test_rules_reproduce_the_oracle(model, spec_from):
    from oracle import oracle as O
    for name, t, v, origin, dx, (ni, nj, nk) in _cases():
        t = np.ascontiguousarray(t, np.uint32)
        v = np.ascontiguousarray(v, np.float32)
        origin = np.ascontiguousarray(origin, np.float32)
        want = O.make_level_set3(t, v, origin, dx, ni, nj, nk, stage=2)
        phi = np.zeros((nk, nj, ni), np.float32)
        ct = np.zeros((nk, nj, ni), np.int32)
        evals = model.sweep_rules_model(t.ctypes.data_as(C.c_void_p), C.c_int64(len(t)), v.ctypes.data_as(C.c_void_p),
                                        origin.ctypes.data_as(C.c_void_p), C.c_float(float(dx)), ni, nj, nk, spec_from,
                                        phi.ctypes.data_as(C.c_void_p), ct.ctypes.data_as(C.c_void_p))
        assert np.array_equal(phi.view(np.uint32).ravel(), want["phi"].view(np.uint32).ravel()), name
        assert np.array_equal(ct.ravel(), want["closest_tri"].ravel()), name
        # the reference evaluates every neighbour that holds a triangle: 7 per node and sweep at most
        assert evals <= 112 * max(ni - 1, 0) * max(nj - 1, 0) * max(nk - 1, 0)
        if name.startswith("table 4,860"):                # the shortcuts are what the model is about: 4 % of the reference's count
            assert evals < 0.1 * 112 * 31 ** 3, evals

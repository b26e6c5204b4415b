"""Randomised parity campaign (GPU box): CUDA vs the CPU oracle on inputs no fixture covers.

    python tools/fuzz_parity.py [--meshes 60] [--poses 300] [--seed 0]

TDF: random triangle soups / jittered tables / degenerate and out-of-cube triangles on random grid sizes,
faithful mode bit-for-bit (phi and closest_tri) and exact mode against the brute-force oracle.
Raycast: random rigid poses (incl. cameras close to / inside the grid) and random depth maps; index map and
occupancy bit-for-bit. Prints one summary line per family; exits non-zero on the first mismatch.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--meshes", type=int, default=60)
    ap.add_argument("--poses", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import torch
    from generate3d_b200 import dfgen, raytracing, _lib, synthetic as S
    from oracle import oracle as O
    rng = np.random.default_rng(a.seed)
    t0 = time.time()

    # ---- TDF
    n_nodes = n_tri = 0
    for m in range(a.meshes):
        kind = m % 4
        if kind == 0:
            nv, nt = int(rng.integers(3, 400)), int(rng.integers(1, 900))
            v = rng.uniform(-0.7, 0.7, (nv, 3)).astype(np.float32)
            t = rng.integers(0, nv, (nt, 3)).astype(np.uint32)
        elif kind == 1:
            v, t = S.table_mesh(int(rng.integers(1, 12)), int(rng.integers(0, 10 ** 6)), scale=rng.uniform(0.3, 1.2),
                                rot_y=rng.uniform(0, 6.28), jitter=10.0 ** rng.uniform(-6, -2))
        elif kind == 2:                                            # slivers, repeated vertices, tiny triangles
            nv = int(rng.integers(6, 100))
            v = rng.uniform(-0.5, 0.5, (nv, 3)).astype(np.float32)
            n3 = min(len(v[::3]), len(v[1::3]))
            v[0:3 * n3:3] = v[1::3][:n3] + np.float32(1e-6) * rng.normal(size=(n3, 3)).astype(np.float32)
            t = rng.integers(0, nv, (int(rng.integers(5, 300)), 3)).astype(np.uint32)
            t[::5, 1] = t[::5, 0]
        else:                                                      # large triangles crossing the whole grid
            v = rng.uniform(-3, 3, (30, 3)).astype(np.float32)
            t = rng.integers(0, 30, (int(rng.integers(1, 12)), 3)).astype(np.uint32)
        ni, nj, nk = (int(x) for x in rng.integers(2, 49, 3))
        if m % 7 == 0:
            ni = nj = nk = 32
        dx = np.float32(rng.uniform(0.5, 1.5) / max(ni, nj, nk))
        origin = (np.float32(-0.5) + rng.uniform(-0.1, 0.1, 3).astype(np.float32))
        want = O.make_level_set3(t, v, origin, dx, ni, nj, nk)
        phi, ct = dfgen.make_level_set3(t, v, origin, dx, ni, nj, nk, return_closest_tri=True)
        ok = np.array_equal(phi.cpu().numpy().view(np.uint32), want["phi"].view(np.uint32)) and \
            np.array_equal(ct.cpu().numpy(), want["closest_tri"])
        if not ok:
            print("TDF faithful MISMATCH at case", m, kind, (ni, nj, nk), len(t))
            sys.exit(1)
        if m % 3 == 0 and len(t) * ni * nj * nk < 4e7:             # exact mode vs brute force (untruncated)
            p = _lib.TdfParams()
            p.ni, p.nj, p.nk = ni, nj, nk
            p.origin[:] = [float(x) for x in origin]
            p.dx, p.out_scale, p.trunc, p.occ_thresh, p.mode, p.exact_band = float(dx), 1.0, float("inf"), 1.0, 1, 1
            r = dfgen.tdf_batch_params(v, np.array([0, len(v)]), t, np.array([0, len(t)]), p, ("sdf", "closest_tri"))
            bf, bt = O.brute_force(t, v, origin, dx, ni, nj, nk)
            got = np.abs(r["sdf"][0].cpu().numpy())
            okx = np.array_equal(got.view(np.uint32), bf.view(np.uint32)) and np.array_equal(r["closest_tri"][0].cpu().numpy(), bt)
            if not okx:
                bad = got.view(np.uint32) != bf.view(np.uint32)
                print("TDF exact MISMATCH at case", m, kind, (ni, nj, nk), len(t), int(bad.sum()), float(np.abs(got - bf).max()))
                sys.exit(1)
        n_nodes += ni * nj * nk
        n_tri += len(t)
    print("tdf: %d random meshes (%d triangles, %d nodes) bit-equal to the oracle, faithful phi+closest_tri and exact vs brute force"
          % (a.meshes, n_tri, n_nodes))

    # ---- raycast
    n_pts = 0
    B = 50
    for b0 in range(0, a.poses, B):
        poses, depths = [], []
        H, W = [(512, 512), (256, 256), (300, 200), (64, 96)][(b0 // B) % 4]
        for q in range(min(B, a.poses - b0)):
            rot = np.linalg.qr(rng.normal(size=(3, 3)))[0]
            p = np.eye(4, dtype=np.float32)
            p[:3, :3] = rot
            p[:3, 3] = rng.uniform(-0.3, 0.3, 3)
            p[2, 3] -= rng.choice([0.0, 0.3, 0.9, 1.4, 2.5])       # from inside the grid to far away
            poses.append(p)
            d = (rng.uniform(size=(H, W)) < 10.0 ** rng.uniform(-4, -0.3)).astype(np.uint8) * 200
            depths.append(d)
        poses, depths = np.stack(poses), np.stack(depths)
        f = 525.0 * W / 512
        r = raytracing.raycast_batch(depths, poses, focal=f, outputs=("occ_f32", "index_map"))
        im = r["index_map"].cpu().numpy()
        occ = r["occ_f32"].cpu().numpy().astype(np.uint8)
        for q in range(poses.shape[0]):
            wim = O.compute_index_mapping(poses[q], 32, f, W / 2, H / 2, 511)
            wocc = O.raycast_depth(depths[q], poses[q], 32, f, W / 2, H / 2, 511)
            if not (np.array_equal(im[q], wim) and np.array_equal(occ[q], wocc)):
                print("RAYCAST MISMATCH at pose", b0 + q, (H, W), int((im[q] != wim).sum()), int((occ[q] != wocc).sum()))
                sys.exit(1)
            n_pts += 33 ** 3
    print("raycast: %d random poses (%d lattice projections) bit-equal to the oracle, index map and occupancy" % (a.poses, n_pts))
    print("fuzz ok in %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()

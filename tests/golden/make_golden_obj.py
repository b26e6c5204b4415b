"""Golden vectors for the OBJ reader: a hand-written OBJ + MTL exercising everything g3d_load_mesh must share with the
reference's loader, parsed by the reference's own vendored tiny_obj_loader.h (compiled in place into
oracle/_ref/libg3d_ref_obj.so) plus the gather loop of LoaderMesh.h:17-50.

    python tests/golden/make_golden_obj.py        (needs /root/reference; writes tests/golden/obj_parse.npz)
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

MTL = "# two materials\nnewmtl red  \nKd 1 0 0\n\nnewmtl blue\nKd 0 0 1\n"


def obj_text(seed=0):
    rng = np.random.default_rng(seed)
    L = ["# golden OBJ for g3d_load_mesh", "mtllib model.mtl missing.mtl", "g first", "usemtl red"]
    # numbers in every spelling tinyobj's own parser meets (tiny_obj_loader.h:505-618)
    L += ["v 0 0 0", "v 1 0 0", "v 1 1 0", "v 0 1 0",
          "v 0.5 0.5 1e0", "v +0.25 -0.25 1E-1", "v 1.5e+1 -2.5E-3 3e2", "v .5 -.5 +.5",          # .5 forms fail -> 0
          "v 1. 2. 3.", "v 1e 2e+ 3e-", "v nan inf -inf", "v 12abc 1.5.3 0x10", "v 1e5x 7 8",
          "v\t0.1\t0.2\t0.3\t0.9 0.8 0.7", "v 0.1234567890123456789 123456789012345678901234567890 1e-40",
          "v 1e38 1e39 -1e39", "v 4.9e-324 1e-400 2", "v 1 2", "v", "v 3.4028235e38 3.4028236e38 1e400",
          "v 0.30000001192092896 0.1 16777217", "   v   5   6   7   "]
    for _ in range(200):
        x = rng.uniform(-1, 1, 3) * 10.0 ** rng.integers(-5, 3, 3)
        L.append(("v %.6f %.9g %.17g", "v %.8e %r %.3f", "v %.4f %.12f %+.7E")[int(rng.integers(3))] % tuple(float(q) for q in x))
    L += ["vn 0 0 1", "vn 0 1 0", "vt 0 0", "vt 1 0", "vt 1 1"]
    # faces: plain, v/vt, v//vn, v/vt/vn, negative (relative), stray spaces, tabs, CR
    L += ["f 1 2 3", "f 1/1 2/2 3/3", "f 1//1 2//1 3//2", "f 1/1/1 2/2/1 3/3/2", "f -1 -2 -3", "f 1/ 2/ 3/1", "f 1//  2//1 3//1",
          "f\t2\t3\t4", "f 1 2 3 4", "f 4 3 2 1", "f 1 2", "f 7", "f  5   6   7   ",
          "usemtl red", "f 2 3 4", "usemtl blue", "f 3 4 1", "usemtl nosuch", "f 1 3 4",
          "g second third", "f 10 11 12 13 14", "f 20 21 22 23 24 25 26 27", "o obj1", "f 5 6 7", "usemtl red", "o obj2",
          "f 1 2 3", "g", "f 30 31 32 33", "f 40 41 42 40 41", "f 1 1 1 1", "f 50 51 52 53 54 55"]
    # a concave polygon and a planar octagon whose corners are real vertices appended at the end
    base = 22 + 200 + 1
    L += ["v 0 0 0", "v 2 0 0", "v 2 2 0", "v 1 0.5 0", "v 0 2 0"]                                  # concave (arrow head)
    L += ["f %d %d %d %d %d" % tuple(range(base, base + 5))]
    ang = np.linspace(0, 2 * np.pi, 9)[:-1]
    for a in ang:
        L.append("v %.6f 0.25 %.6f" % (np.cos(a), np.sin(a)))
    L += ["f " + " ".join(str(base + 5 + k) for k in range(8)), "f " + " ".join(str(-(k + 1)) for k in range(8))]
    L += ["v 0 0 0", "v 1 0 0", "v 2 0 0", "v 3 0 0", "f -1 -2 -3 -4"]                              # collinear quad
    return "\r\n".join(L[:60]) + "\r" + "\n".join(L[60:]) + "\n"


def main():
    from oracle import oracle as O
    d = tempfile.mkdtemp()
    text = obj_text()
    with open(os.path.join(d, "model.obj"), "w", newline="") as f:
        f.write(text)
    with open(os.path.join(d, "model.mtl"), "w") as f:
        f.write(MTL)
    v, fi = O.ref_load_obj(os.path.join(d, "model.obj"))
    # inputs tinyobj refuses as a whole (LoaderMesh.h:24 then asserts): index 0, a non-number, an empty vt / vn field
    bad_texts = ["v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\nf 0 1 2\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf a b c\n",
                 "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1/ 2/ 3/\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1//0 2//1 3//1\n"]
    for bad_text in bad_texts:
        with open(os.path.join(d, "bad.obj"), "w") as f:
            f.write(bad_text)
        assert O.ref_load_obj(os.path.join(d, "bad.obj")) is None, bad_text
    bad_text = "\0".join(bad_texts)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "obj_parse.npz"), obj=np.frombuffer(text.encode(), np.uint8),
                        mtl=np.frombuffer(MTL.encode(), np.uint8), verts=v, corners=fi,
                        bad_obj=np.frombuffer(bad_text.encode(), np.uint8))
    print("verts", v.shape, "corners", fi.shape, "->", fi.shape[0] // 3, "triangles")


if __name__ == "__main__":
    main()

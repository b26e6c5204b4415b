"""Golden fixture for the per-pixel ray march, produced by the REFERENCE's own
raytracing_approach1.raycast + VoxelGrid (build container only). A 96x96 view keeps the reference's
pure-Python double loop to about a minute."""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from make_golden import import_reference, REF  # noqa: E402


def main():
    from PIL import Image
    import_reference()
    cwd = os.getcwd()
    os.chdir(REF)
    import raytracing_approach1 as ra
    os.chdir(cwd)
    rng = np.random.default_rng(11)
    H = W = 96
    depth = np.zeros((H, W), np.uint8)
    depth[20:70, 15:80] = 120
    depth[rng.integers(0, H, 200), rng.integers(0, W, 200)] = 50
    depth[40:50, 30:40] = 0
    color = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    tmp = tempfile.mkdtemp()
    Image.fromarray(depth, mode="L").save(os.path.join(tmp, "d.png"))
    Image.fromarray(color, mode="RGB").save(os.path.join(tmp, "c.png"))
    cam_json = {"intrinsic": {"cx": 48, "cy": 48, "fx": 98.4375, "fy": 98.4375, "z_near": 0.5, "z_far": 1.5},
                "pose": [float(x) for x in np.eye(4).reshape(-1)]}
    with open(os.path.join(tmp, "cam.json"), "w") as f:
        json.dump(cam_json, f)
    cam = ra.load_camera(os.path.join(tmp, "cam.json"))
    grid = ra.create_voxel_grid(cam)
    ra.raycast(grid, cam, os.path.join(tmp, "c.png"), os.path.join(tmp, "d.png"))
    grid.save_vox(os.path.join(tmp, "o.txt"))
    with open(os.path.join(tmp, "o.txt")) as f:
        text = f.read()
    np.savez_compressed(os.path.join(HERE, "raymarch.npz"), depth=depth, color=color, cam=json.dumps(cam_json),
                        occ=grid.occ_grid.astype(np.uint8), rgb=grid.color_grid.astype(np.uint8),
                        txt=np.frombuffer(text.encode(), np.uint8))
    print("raymarch: occupied", int(grid.occ_grid.sum()))


if __name__ == "__main__":
    main()

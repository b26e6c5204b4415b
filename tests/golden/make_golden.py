"""Generate tests/golden/*.npz by running the REFERENCE's own Python code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference's raytracing.py, dataset_loader.py and losses.py are imported
unmodified (four unused third-party imports are stubbed, and the two config
attributes raytracing.py:31-32 reads but config.py does not define are injected,
SURVEY.md App. A.2) and executed under CPU torch. Their outputs on seeded
synthetic inputs are stored as small fixtures; the inputs are stored next to
them so that no test needs the reference at run time.

Fixtures
  raycast_512.npz   reference-native 512x512 / focal 525, 3 views
  raycast_256.npz   config[1] shape: 256x256 / focal 262.5, 20 views
  raycast_rand.npz  8 random rigid poses, random blob depth maps (stress for
                    the rounding / association of the projection)
  loader.npz        load_pose / load_df / load_occ / apply_log_transform on a
                    .df and .txt produced by the CPU oracle
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src/scripts"


def import_reference():
    sys.path[:0] = [REF, REF + "/Network"]
    for m in ("trimesh", "imageio", "plyfile", "tinyobjloader"):
        sys.modules.setdefault(m, types.ModuleType(m))
    import config
    config.render_img_width = config.render_img_height = 512
    cwd = os.getcwd()
    os.chdir(REF)
    import raytracing
    import dataset_loader
    import losses
    os.chdir(cwd)
    return config, raytracing, dataset_loader, losses


def run_view(raytracing, tmp, depth, pose_payload):
    """One reference raycast_depth call through its real file interface."""
    from PIL import Image
    dpng, pjson, txt = (os.path.join(tmp, n) for n in ("depth.png", "pose.json", "out.txt"))
    Image.fromarray(depth, mode="L").save(dpng)
    with open(pjson, "w") as f:
        json.dump({"pose": pose_payload}, f)
    raytracing.raycast_depth(txt, pjson, dpng)
    occ = np.zeros((32, 32, 32), np.uint8)                      # [z][y][x]
    with open(txt) as f:
        text = f.read()
    for line in text.splitlines():
        x, y, z = (int(t) for t in line.split()[:3])
        occ[z, y, x] = 1
    return occ, text


def index_map_of(raytracing, dataset_loader, tmp, pose_payload):
    import torch
    pjson = os.path.join(tmp, "pose.json")
    with open(pjson, "w") as f:
        json.dump({"pose": pose_payload}, f)
    pose = dataset_loader.load_pose(pjson)
    g2w = torch.inverse(raytracing.make_world_to_grid())
    im = raytracing.compute_index_mapping(pose, g2w, torch.device("cpu"))
    return pose.numpy(), im.numpy().reshape(4, -1)


def main():
    sys.path.insert(0, ROOT)
    from generate3d_b200 import synthetic as S
    from oracle import oracle as O
    config, raytracing, dataset_loader, losses = import_reference()
    tmp = tempfile.mkdtemp()

    def raycast_set(name, payloads, depths, W, focal):
        config.render_img_width = config.render_img_height = W
        config.focal = focal
        poses, ims, occs, texts = [], [], [], []
        for payload, depth in zip(payloads, depths):
            pose, im = index_map_of(raytracing, dataset_loader, tmp, payload)
            occ, text = run_view(raytracing, tmp, depth, payload)
            mine = O.compute_index_mapping(pose, 32, focal, W / 2, W / 2, 511)
            mocc = O.raycast_depth(depth, pose, 32, focal, W / 2, W / 2, 511)
            print(name, "index_map mismatches vs restatement:", int((mine != im).sum()),
                  "occ mismatches:", int((mocc != occ).sum()), "occupied:", int(occ.sum()))
            assert O.occ_txt_rows(occ) == text
            poses.append(pose)
            ims.append(im.astype(np.int16))
            occs.append(np.packbits(occ.reshape(-1), bitorder="little"))
            texts.append(text)
        np.savez_compressed(os.path.join(HERE, name), pose=np.stack(poses), depth=np.stack(depths),
                            index_map=np.stack(ims), occ_bits=np.stack(occs),
                            focal=np.float32(focal), first_txt=np.frombuffer(texts[0].encode(), np.uint8))
        config.render_img_width = config.render_img_height = 512
        config.focal = 525.0

    raycast_set("raycast_512.npz", [S.pose_json_payload(k) for k in (0, 5, 13)],
                [S.depth_view(k, 512, 512) for k in (0, 5, 13)], 512, 525.0)
    raycast_set("raycast_256.npz", [S.pose_json_payload(k) for k in range(20)],
                [S.depth_view(k, 256, 256) for k in range(20)], 256, 262.5)
    rng = np.random.default_rng(7)
    payloads, depths = [], []
    for q in range(8):
        rot = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        pose = np.eye(4); pose[:3, :3] = rot
        payloads.append([float(x) for x in pose.T.reshape(-1)])
        d = np.zeros((512, 512), np.uint8)
        for _ in range(12):
            cy, cx, r = rng.integers(60, 450, 2).tolist() + [int(rng.integers(3, 40))]
            d[max(cy - r, 0):cy + r, max(cx - r, 0):cx + r] = rng.integers(1, 255)
        d[rng.integers(0, 512, 300), rng.integers(0, 512, 300)] = 200    # isolated pixels
        depths.append(d)
    raycast_set("raycast_rand.npz", payloads, depths, 512, 525.0)

    # ---- loader fixtures: .df / .txt written by the CPU oracle, read by the reference
    v, t = S.table_mesh(9, seed=1)
    g = O.dfgen(t, v, 32)
    dfp = os.path.join(tmp, "m__0__.df")
    O.write_df(dfp, g["sdf"], g["grid2world"])
    txtp = os.path.join(tmp, "m__0__.txt")
    with open(txtp, "w") as f:
        f.write(O.gt_occ_txt_rows(g["sdf"]))
    df3 = dataset_loader.load_df(dfp, 3.0).numpy()
    occ = dataset_loader.load_occ(txtp).numpy()
    import torch
    tdflog = losses.apply_log_transform(torch.from_numpy(df3)).numpy()
    pj = os.path.join(tmp, "p.json")
    with open(pj, "w") as f:
        json.dump({"pose": S.pose_json_payload(3)}, f)
    pose = dataset_loader.load_pose(pj).numpy()
    assert np.array_equal(df3[0], O.load_df_clamp(g["sdf"], 3.0))
    assert np.array_equal(occ[0].astype(np.uint8), O.occupancy(g["sdf"]))
    print("tdflog max |ref - restatement|:", float(np.abs(tdflog[0] - O.apply_log_transform(df3[0])).max()))
    print("load_pose == synthetic.view_pose:", np.array_equal(pose, S.view_pose(3)))
    with open(dfp, "rb") as f:
        df_bytes = np.frombuffer(f.read(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "loader.npz"), verts=v, tris=t, df_file=df_bytes,
                        tdf3=df3, occ_gt=occ.astype(np.uint8), tdflog3=tdflog, pose3=pose,
                        pose3_payload=np.array(S.pose_json_payload(3)))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

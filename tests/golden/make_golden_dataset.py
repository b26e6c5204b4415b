"""Golden fixture for the reference-signature data-set path: a three-model data-set directory in the reference's
layout (parameters.json, <raytraced>/<synset>/<id>.txt, <voxelized>/<synset>/<id>__0__.df / __0__.txt /
_occ_df.npy) and the items the REFERENCE's own DatasetLoad(data_list, truncation) (Network/dataset_loader.py:155-189)
returns for it under CPU torch. Build container only (needs /root/reference):

    python tests/golden/make_golden_dataset.py      -> tests/golden/dataset_dir.npz

Files are produced by the reference's own code where it runs here (raycast_depth, occ_to_df.py's occ_to_df) and by
the CPU oracle for the dfgen / vox2mesh outputs (their binaries need Eigen).
"""
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
from make_golden_next import legacy_integer_division  # noqa: E402

SYNSET = "04379243"


def main():
    import torch
    from PIL import Image
    from generate3d_b200 import synthetic as S
    from oracle import oracle as O
    config, raytracing, dataset_loader, losses = import_reference()
    cwd = os.getcwd()
    os.chdir(REF)
    import occ_to_df as occ_to_df_script
    os.chdir(cwd)
    tmp = tempfile.mkdtemp()
    ray, vox, work = (os.path.join(tmp, d) for d in ("raytraced", "voxelized", "work"))
    for d in (os.path.join(ray, SYNSET), os.path.join(vox, SYNSET), work):
        os.makedirs(d)
    params = {"shapenet_raytraced": ray + "/", "shapenet_voxelized": vox + "/"}
    with open(os.path.join(tmp, "parameters.json"), "w") as f:
        json.dump(params, f)
    names = ["model_a", "model_b", "model_c"]
    files = {}
    for k, name in enumerate(names):
        v, t = S.table_mesh(4 + k, seed=40 + k, scale=0.8 + 0.1 * k, rot_y=0.5 * k)
        g = O.dfgen(t, v, 32)
        O.write_df(os.path.join(vox, SYNSET, name + "__0__.df"), g["sdf"], g["grid2world"])
        with open(os.path.join(vox, SYNSET, name + "__0__.txt"), "w") as f:
            f.write(O.gt_occ_txt_rows(g["sdf"]))
        occ_gt = dataset_loader.load_occ(os.path.join(vox, SYNSET, name + "__0__.txt"))
        with legacy_integer_division():
            df = occ_to_df_script.occ_to_df(torch.transpose(occ_gt[0], 0, 2).clone(), 3.0)          # occ_to_df.py:70-74
        np.save(os.path.join(vox, SYNSET, name + "_occ_df"), df.cpu().numpy())
        dep = S.depth_view(3 * k, 512, 512)
        Image.fromarray(dep, mode="L").save(os.path.join(tmp, "d.png"))
        with open(os.path.join(tmp, "p.json"), "w") as f:
            json.dump({"pose": S.pose_json_payload(3 * k)}, f)
        raytracing.raycast_depth(os.path.join(ray, SYNSET, name + ".txt"), os.path.join(tmp, "p.json"), os.path.join(tmp, "d.png"))
        for rel in ("raytraced/%s/%s.txt" % (SYNSET, name), "voxelized/%s/%s__0__.df" % (SYNSET, name),
                    "voxelized/%s/%s__0__.txt" % (SYNSET, name), "voxelized/%s/%s_occ_df.npy" % (SYNSET, name)):
            with open(os.path.join(tmp, rel), "rb") as f:
                files[rel] = np.frombuffer(f.read(), np.uint8)
    data_list = [{"synset_id": SYNSET, "model_id": n} for n in names]
    os.chdir(work)                                                    # DatasetLoad opens "../parameters.json"
    ds = dataset_loader.DatasetLoad(data_list, 3.0)
    items = [ds[i] for i in range(len(ds))]
    os.chdir(cwd)
    out = {"names": np.array(names), "synset": np.array(SYNSET), "file_names": np.array(sorted(files))}
    for i, rel in enumerate(sorted(files)):
        out["file_%d" % i] = files[rel]
    for key in ("occ_grid", "occ_gt", "occ_df_gt", "df_gt"):
        out[key] = np.stack([it[key].numpy() for it in items]).astype(np.float32)
        assert out[key].shape == (3, 1, 32, 32, 32), out[key].shape
    assert [it["name"] for it in items] == names
    np.savez_compressed(os.path.join(HERE, "dataset_dir.npz"), **out)
    print({k: (v.shape, float(v.sum())) for k, v in out.items() if k in ("occ_grid", "occ_gt", "occ_df_gt", "df_gt")})


if __name__ == "__main__":
    main()

"""The reference-signature data-set path: DatasetLoad(data_list, truncation) over a
directory in the reference's layout, against the items the reference's own DatasetLoad returned for the same files
(tests/golden/dataset_dir.npz, made by tests/golden/make_golden_dataset.py)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _materialise(z, root):
    for i, rel in enumerate(z["file_names"]):
        path = os.path.join(root, str(rel))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            f.write(bytes(z["file_%d" % i]))
    params = {"shapenet_raytraced": os.path.join(root, "raytraced") + "/", "shapenet_voxelized": os.path.join(root, "voxelized") + "/"}
    with open(os.path.join(root, "parameters.json"), "w") as f:
        json.dump(params, f)
    return params


def test_datasetload_constructor_path_equals_reference_items(g3d, golden_dir, tmp_path, monkeypatch):
    from generate3d_b200 import dataset_loader
    z = np.load(os.path.join(golden_dir, "dataset_dir.npz"))
    params = _materialise(z, str(tmp_path))
    names = [str(n) for n in z["names"]]
    data_list = [{"synset_id": str(z["synset"]), "model_id": n} for n in names]
    # (1) exactly as the reference is used: cwd one level below parameters.json (dataset_loader.py:177)
    work = tmp_path / "work"
    work.mkdir()
    monkeypatch.chdir(work)
    ds = dataset_loader.DatasetLoad(data_list, 3.0)
    assert len(ds) == 3
    for i, name in enumerate(names):
        item = ds[i]
        assert set(item) == {"df_gt", "name", "occ_df_gt", "occ_gt", "occ_grid"} and item["name"] == name
        for k in ("occ_grid", "occ_gt", "occ_df_gt", "df_gt"):
            assert item[k].is_cuda and item[k].dtype == torch.float32 and tuple(item[k].shape) == (1, 32, 32, 32)
            assert np.array_equal(item[k].cpu().numpy().view(np.uint32), z[k][i].view(np.uint32)), (name, k)
    # (2) through a DataLoader, as the trainers consume it (trainer_df.py:81, with num_workers=0)
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=3, shuffle=False, num_workers=0)))
    assert batch["name"] == names
    for k in ("occ_grid", "occ_gt", "occ_df_gt", "df_gt"):
        assert tuple(batch[k].shape) == (3, 1, 32, 32, 32) and batch[k].is_cuda
        assert np.array_equal(batch[k].cpu().numpy(), z[k])
    # (3) n_max_samples and explicit params
    ds2 = dataset_loader.DatasetLoad(data_list, 3.0, n_max_samples=2, params=params)
    assert len(ds2) == 2
    with pytest.raises(IndexError):
        ds2[2]


def test_datasetload_missing_file_fails_that_item_only(g3d, golden_dir, tmp_path):
    """The reference raises from __getitem__ of the item whose file is missing (np.load / np.loadtxt) and serves the
    others; it never drops a key."""
    from generate3d_b200 import dataset_loader
    z = np.load(os.path.join(golden_dir, "dataset_dir.npz"))
    params = _materialise(z, str(tmp_path))
    names = [str(n) for n in z["names"]]
    os.remove(os.path.join(params["shapenet_voxelized"], str(z["synset"]), names[1] + "_occ_df.npy"))
    ds = dataset_loader.DatasetLoad([{"synset_id": str(z["synset"]), "model_id": n} for n in names], 3.0, params=params)
    assert set(ds[0]) == {"name", "occ_grid", "occ_gt", "occ_df_gt", "df_gt"} and set(ds[2]) == set(ds[0])
    with pytest.raises(FileNotFoundError) as ei:
        ds[1]
    assert "_occ_df.npy" in str(ei.value)


def test_datasetload_refuses_worker_processes_with_a_clear_message(g3d, golden_dir, tmp_path):
    from generate3d_b200 import dataset_loader
    z = np.load(os.path.join(golden_dir, "dataset_dir.npz"))
    params = _materialise(z, str(tmp_path))
    ds = dataset_loader.DatasetLoad([{"synset_id": str(z["synset"]), "model_id": str(n)} for n in z["names"]], 3.0, params=params)
    with pytest.raises(RuntimeError) as ei:
        next(iter(torch.utils.data.DataLoader(ds, batch_size=1, num_workers=1)))
    assert "num_workers=0" in str(ei.value)


def test_datasetload_from_tensors_always_has_five_keys(g3d, golden_dir):
    """from_tensors without an occ_df_gt computes it on the GPU: equal to what the reference derives from the same
    GT occupancy through occ_to_df.py + load_occ_df."""
    from generate3d_b200 import dataset_loader
    z = np.load(os.path.join(golden_dir, "dataset_dir.npz"))
    dev = torch.device("cuda", 0)
    g = {k: torch.from_numpy(z[k][:, 0]).to(dev) for k in ("occ_grid", "occ_gt", "df_gt")}
    ds = dataset_loader.DatasetLoad.from_tensors([str(n) for n in z["names"]], g["occ_grid"], g["occ_gt"], g["df_gt"])
    for i in range(3):
        item = ds[i]
        assert set(item) == {"name", "occ_grid", "occ_gt", "occ_df_gt", "df_gt"}
        assert np.array_equal(item["occ_df_gt"].cpu().numpy(), z["occ_df_gt"][i])

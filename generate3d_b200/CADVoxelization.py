"""Drop-in for src/scripts/CADVoxelization.py (README step 5): voxelise every ShapeNet model to a 32^3 distance
field. The reference spawns `../dfgen/main` and `../vox2mesh/main` once per model (CADVoxelization.py:15-39);
here OBJ files are parsed on host threads (g3d_load_mesh releases the GIL), meshes are batched through one
g3d_tdf_batch call per chunk, and each model still gets the same two files:

    <shapenet_voxelized>/<synset>/<model>__0__.df    byte-identical to dfgen's output (faithful mode)
    <shapenet_voxelized>/<synset>/<model>__0__.txt   the GT occupancy rows vox2mesh writes (|sdf| <= 1)

Models are sharded over ranks with no collective (generate3d_b200/shard.py); existing outputs are skipped like
raytracing.py:159-161 does, so an interrupted run resumes. Per-model failures are collected, not fatal
(CADVoxelization.py:28-29 swallows CalledProcessError the same way).

    python -m generate3d_b200.CADVoxelization [parameters.json] [--chunk 512] [--mode faithful|exact] [--ply]
"""
import glob
import json
import os
import pathlib
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import dfgen, shard
from .config import vox_dim


def list_models(params):
    """(obj_path, catid, id) for params['shapenet']/**/*/models/model_normalized.obj (CADVoxelization.py:15-17)."""
    out = []
    for f in sorted(glob.glob(params["shapenet"] + "/**/*/models/model_normalized.obj", recursive=True)):
        parts = f.replace("\\", "/").split("/")
        out.append((f, parts[-4], parts[-3]))
    return out


def voxelize_models(models, out_root, dim=vox_dim, chunk=512, mode="faithful", write_txt=True, threads=None,
                    skip_existing=True, keep=None, write_ply=False, tri_budget=24 << 20):
    """models: [(obj_path, catid, model_id)]. Returns (n_done, failed list). `keep` (optional dict) receives the
    GPU-resident 'tdf'/'occ_f32' tensors keyed by model id, for handing straight to DatasetLoad.from_tensors."""
    threads = threads or min(32, os.cpu_count() or 1)
    _, g2w = dfgen.dfgen_geometry(dim, 0, 1)
    failed, done = [], 0
    todo = []
    for f, cat, mid in models:
        outdir = os.path.join(out_root, cat)
        df = os.path.join(outdir, mid + "__0__.df")
        if skip_existing and os.path.exists(df) and (not write_txt or os.path.exists(df[:-3] + ".txt")):
            continue
        todo.append((f, cat, mid, outdir, df))

    def load(item):
        try:
            v, t = dfgen.load_mesh(item[0])
            if len(v) == 0 or len(t) == 0:
                raise ValueError("Error loading mesh.")          # LoaderMesh.h:83
            return v, t
        except Exception as e:                                   # noqa: BLE001 -- per-model isolation
            return e

    with ThreadPoolExecutor(threads) as pool:
        for c0 in range(0, len(todo), chunk):
            items = todo[c0:c0 + chunk]
            loaded = list(pool.map(load, items))
            good = [(it, m) for it, m in zip(items, loaded) if not isinstance(m, Exception)]
            failed += [(it[2], str(m)) for it, m in zip(items, loaded) if isinstance(m, Exception)]
            # heavy models: keep one g3d_tdf_batch call under `tri_budget` triangles (about 80 B of scratch each)
            parts, cur, cur_t = [], [], 0
            for g in good:
                if cur and cur_t + len(g[1][1]) > tri_budget:
                    parts.append(cur)
                    cur, cur_t = [], 0
                cur.append(g)
                cur_t += len(g[1][1])
            if cur:
                parts.append(cur)
            for part in parts:
                done += _voxelize_part(part, pool, dim, mode, write_txt, write_ply, keep, g2w, failed)
    return done, failed


def _solve(good, dim, mode, outs):
    voff, toff = [0], [0]
    for _, (v, t) in good:
        voff.append(voff[-1] + len(v))
        toff.append(toff[-1] + len(t))
    V = np.concatenate([m[0] for _, m in good])
    T = np.concatenate([m[1] for _, m in good])
    return dfgen.tdf_batch(V, np.asarray(voff, np.int64), T, np.asarray(toff, np.int64), dim=dim, mode=mode, outputs=outs)


def _voxelize_part(good, pool, dim, mode, write_txt, write_ply, keep, g2w, failed):
    """One g3d_tdf_batch call for `good` (a list of (item, (verts, tris))). A failure of the call as a whole (device
    out of memory, an invalid mesh the library refuses) does not take the other models down: the part is solved
    again one model at a time and only the offenders are reported -- the per-model isolation the reference gets from
    its try/except around each subprocess (CADVoxelization.py:24-29)."""
    from ._lib import G3DError
    outs = ("sdf", "occ_f32", "status") + (("tdf",) if keep is not None else ())
    try:
        results = [(good, _solve(good, dim, mode, outs))]
    except (G3DError, RuntimeError) as e:                        # RuntimeError: torch's out-of-memory
        if len(good) == 1:
            failed.append((good[0][0][2], str(e)))
            return 0
        results = []
        for g in good:
            try:
                results.append(([g], _solve([g], dim, mode, outs)))
            except (G3DError, RuntimeError) as e1:
                failed.append((g[0][2], str(e1)))
    n = 0
    for good, r in results:
        sdf = r["sdf"].cpu().numpy()
        occ = r["occ_f32"].cpu().numpy()
        status = r["status"].cpu().numpy()

        def write(k):
            (f, cat, mid, outdir, df) = good[k][0]
            pathlib.Path(outdir).mkdir(parents=True, exist_ok=True)
            dfgen.write_df(df, sdf[k], g2w)
            if write_ply:                                        # CADVoxelization.py:33-37: vox2mesh writes .txt and .ply
                dfgen.vox2mesh(df, df[:-3] + ".ply", is_unitless=True)
            elif write_txt:
                with open(df[:-3] + ".txt", "w") as fh:
                    fh.write(dfgen.occ_rows_gt(occ[k]))
            return mid if status[k] else None
        for bad in pool.map(write, range(len(good))):
            if bad:
                failed.append((bad, "mesh status flag (out-of-range face index)"))
        if keep is not None:
            for k, (it, _) in enumerate(good):
                keep[it[2]] = {"tdf": r["tdf"][k], "occ_f32": r["occ_f32"][k]}
        n += len(good)
    return n


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    pfile = next((a for a in argv if not a.startswith("--")), "./parameters.json")
    chunk = int(argv[argv.index("--chunk") + 1]) if "--chunk" in argv else 512
    mode = argv[argv.index("--mode") + 1] if "--mode" in argv else "faithful"
    ply = "--ply" in argv
    with open(pfile) as f:
        params = json.load(f)                                   # CADVoxelization.py:12
    rank, world, local = shard.init_process_group()
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    models = list_models(params)
    lo, hi = shard.shard_range(len(models), rank, world)
    done, failed = voxelize_models(models[lo:hi], params["shapenet_voxelized"], chunk=chunk, mode=mode, write_ply=ply)
    print("rank %d: voxelised %d of %d models, %d failed" % (rank, done, hi - lo, len(failed)))
    if failed:
        with open(os.path.join(params["shapenet_voxelized"], "failed_cases_rank%d.json" % rank), "w") as f:
            json.dump(failed, f)


if __name__ == "__main__":
    main()

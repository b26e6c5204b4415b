cd $GRAFT_REPO_ROOT
python tools/fuzz_parity.py --meshes 100 --poses 0 --seed 121 2>&1 | tail -1
G3D_SWEEP_SPEC=-1 G3D_SWEEP_SPLIT=3 G3D_INIT_TILED=1 G3D_SWEEP_THREADS=128 python tools/fuzz_parity.py --meshes 80 --poses 0 --seed 126 2>&1 | tail -1
G3D_SWEEP_BITMAP=0 G3D_SWEEP_THREADS=256 python tools/fuzz_parity.py --meshes 60 --poses 0 --seed 127 2>&1 | tail -1
python tools/faithful_time.py 1024 2>&1 | tail -1
python -m pytest tests/test_gpu_tdf.py -x -q -m gpu -k "config3 or config0 or ragged or noncubic" 2>&1 | tail -3

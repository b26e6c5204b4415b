cd $GRAFT_REPO_ROOT
python bench.py --no-cpu-baseline --no-other-mode --no-next-rows --no-configs > gpurun_out/e2e_test.json 2> gpurun_out/e2e_test.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/e2e_test.json').read().strip().splitlines()[-1])
print("value ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "sync ms", d["e2e"]["sync"]["ms_per_step"])
PY
python -m pytest tests/test_gpu_tdf.py -x -q -m gpu -k "host" 2>&1 | tail -2

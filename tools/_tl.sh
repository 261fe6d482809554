cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_adapter.py -x -q -m gpu -k "upload or fetch" 2>&1 | tail -2
timeout 400 python bench.py --no-extras --no-cpu-baseline --steps 5 > gpurun_out/bench_quick.txt 2> gpurun_out/bench_quick.err; echo "quick rc=$?"
python - <<'PY'
import json
for f in ("bench_quick",):
    d = json.loads([l for l in open(f"gpurun_out/{f}.txt").read().splitlines() if l.startswith("{")][-1])
    print(f, d["value"], d["ms_per_step"], "e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "ms_each_step", "videos")})
PY

cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
for w in 1 2 3; do
  echo "workers $w"; A3D_PIPELINE_WORKERS=$w timeout 300 python tools/e2e_profile.py c3 6 2>&1 | head -4
done | tee gpurun_out/r2_e2e_workers.txt
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_w2.txt 2>/dev/null
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_w2.txt").read().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "ms_each_step", "videos", "device_passes_per_step_rank0")})
PY

cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for om in 0 1; do echo "== c3_mini out_mode=$om"; AB_ITERS=10 AB_OUT_MODE=$om timeout 300 python tools/project_ab.py c3_mini 2>&1 | tail -2; done 2>&1 | tee gpurun_out/r2_ab13.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c13.txt 2>/dev/null
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_c13.txt").read().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernels_ms"], "frac", d["roofline"]["frac"])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "ms_each_step", "videos", "device_passes_per_step_rank0")})
print("c2", d["extras"]["c2"]["ms_per_step"], d["extras"]["c2"]["roofline"]["kernels_ms"])
PY

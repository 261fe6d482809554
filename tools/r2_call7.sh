cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
for o in 0 1; do
  A3D_SCORE_ORDER=$o timeout 600 python bench.py --workload c3_shard --no-cpu-baseline --e2e-videos 1 --no-extras > gpurun_out/bench_shard_o$o.txt 2>/dev/null
done
A3D_SCORE_ORDER=1 timeout 600 python bench.py --no-cpu-baseline --e2e-videos 1 --no-extras > gpurun_out/bench_c3_o1.txt 2>/dev/null
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c3_def.txt 2>/dev/null
python - <<'PY'
import json
for f in ("shard_o0", "shard_o1", "c3_o1", "c3_def"):
    d = json.loads([l for l in open(f"gpurun_out/bench_{f}.txt").read().splitlines() if l.startswith("{")][-1])
    print(f, {k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernels_ms"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("ms_each_step"))
PY
timeout 200 python tools/score_ab.py c3_shard ldg mma > gpurun_out/score_ab_c3_shard.txt 2>&1; tail -3 gpurun_out/score_ab_c3_shard.txt

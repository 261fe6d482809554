cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_s2.txt 2> gpurun_out/bench_n1_s2.err; echo "rc=$?"; tail -2 gpurun_out/bench_n1_s2.err
timeout 600 python bench.py --no-cpu-baseline --streams 1 --e2e-videos 1 --no-extras > gpurun_out/bench_n1_s1.txt 2> gpurun_out/bench_n1_s1.err; echo "rc=$?"
timeout 600 python bench.py --workload c3_shard --no-cpu-baseline --e2e-videos 1 --no-extras > gpurun_out/bench_shard_s2.txt 2>/dev/null
timeout 600 python bench.py --workload c3_shard --no-cpu-baseline --e2e-videos 1 --no-extras --streams 1 > gpurun_out/bench_shard_s1.txt 2>/dev/null
python - <<'PY'
import json
for f in ("n1_s2", "n1_s1", "shard_s2", "shard_s1"):
    d = json.loads([l for l in open(f"gpurun_out/bench_{f}.txt").read().splitlines() if l.startswith("{")][-1])
    print(f, {k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernels_ms"], d["clocks"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("ms_each_step"))
PY

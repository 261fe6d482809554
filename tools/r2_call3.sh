#!/bin/bash
# GPU call 3: full parity suite, smoke, bench lines (c3 headline, reference arm, c4 shard + translations), ncu captures
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.txt
timeout 600 python bench.py > gpurun_out/bench_default.txt 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_default.txt").read().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernels_ms"], "frac", d["roofline"]["frac"])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "ms_each_step", "videos", "device_passes_per_step_rank0")})
print("c2", d["extras"]["c2"]["ms_per_step"], d["extras"]["c2"]["roofline"]["kernels_ms"])
PY
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.txt 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
for wl in c4_shard c4_trans; do
  timeout 600 python bench.py --workload $wl --steps 10 --e2e-videos 1 --no-extras --no-cpu-baseline > gpurun_out/bench_$wl.txt 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_$wl.txt").read().splitlines() if l.startswith("{")][-1])
print("$wl", {k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernels_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
done
for spec in "c3_shard k_unproject" "c3_shard k_project" "c3_shard k_score_mma" "c3_shard k_finalize" "c2 k_unproject" "c2 k_project" "c2 k_score" "c2 k_finalize" "c3 k_project" "c3 k_score_mma" "c4_shard k_project" "c4_shard k_score_mma"; do
  set -- $spec
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f \
    -o gpurun_out/prof_$1_$2 python tools/profile_pass.py --workload $1 --passes 2 > gpurun_out/ncu_$1_$2.log 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --workload c3_shard --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-videos 1 > gpurun_out/launches_bench.log 2>&1
python tools/refresh_profiles.py r2 > gpurun_out/refresh.log 2>&1; tail -2 gpurun_out/refresh.log
mkdir -p gpurun_out/profiles_new; cp profiles/r2_* profiles/traffic.json profiles/pipes.json gpurun_out/profiles_new/ 2>/dev/null
find gpurun_out -name '*.ncu-rep' ! -name 'prof_c3_shard_k_project.ncu-rep' -delete; du -sh gpurun_out | tail -1

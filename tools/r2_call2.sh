#!/bin/bash
# GPU call: parity tests of the new kernels, old-vs-new projection A/B, bench line, pipe micro-benchmark
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest2.log
tail -5 gpurun_out/r2_gputest2.log
for wl in c3_mini c2; do
  (cd _ab/old && AB_ITERS=10 timeout 300 python tools/project_ab.py $wl) > gpurun_out/r2_ab_old_$wl.txt 2>&1
  AB_ITERS=10 AB_OUT_MODE=0 timeout 300 python tools/project_ab.py $wl > gpurun_out/r2_ab_new_full_$wl.txt 2>&1
  AB_ITERS=10 AB_OUT_MODE=1 timeout 300 python tools/project_ab.py $wl > gpurun_out/r2_ab_new_rows_$wl.txt 2>&1
  tail -4 gpurun_out/r2_ab_old_$wl.txt gpurun_out/r2_ab_new_full_$wl.txt gpurun_out/r2_ab_new_rows_$wl.txt
done
timeout 600 python bench.py > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1b.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernels_ms"], "frac", d["roofline"]["frac"])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "ms_each_step", "videos", "device_passes_per_step_rank0")})
print("c2", d["extras"]["c2"]["ms_per_step"], d["extras"]["c2"]["roofline"]["kernels_ms"])
PY
tail -3 gpurun_out/r2_bench_n1b.err
tools/_build/pipes_bench > gpurun_out/r2_pipes_bench.txt 2>&1; cat gpurun_out/r2_pipes_bench.txt

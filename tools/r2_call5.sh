#!/bin/bash
# GPU call 5: A/B of the experimental decision code, parity suite with both builds, bench (prefetched table preparation)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for wl in c3_mini c2; do
  echo "== $wl old"; (cd _ab/old && AB_ITERS=10 timeout 300 python tools/project_ab.py $wl) 2>&1 | tail -2
  for om in 0 1; do
    echo "== $wl new out_mode=$om"; AB_ITERS=10 AB_OUT_MODE=$om timeout 300 python tools/project_ab.py $wl 2>&1 | tail -2
    echo "== $wl exp2 out_mode=$om"; A3D_LIB=$PWD/tools/_build/liba3d_exp2.so AB_ITERS=10 AB_OUT_MODE=$om timeout 300 python tools/project_ab.py $wl 2>&1 | tail -2
  done
done > gpurun_out/r2_ab5.txt 2>&1; cat gpurun_out/r2_ab5.txt
A3D_LIB=$PWD/tools/_build/liba3d_exp2.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_exp2.txt 2>&1; echo "pytest exp2 rc=$?"; tail -3 gpurun_out/pytest_gpu_exp2.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/bench_default.txt 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_default.txt").read().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernels_ms"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "ms_each_step", "videos", "device_passes_per_step_rank0")})
print("c2", d["extras"]["c2"]["ms_per_step"], d["extras"]["c2"]["roofline"]["kernels_ms"])
PY
A3D_LIB=$PWD/tools/_build/liba3d_exp2.so timeout 600 python bench.py --no-cpu-baseline --e2e-videos 1 > gpurun_out/bench_exp2.txt 2> gpurun_out/bench_exp2.err; echo "bench exp2 rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_exp2.txt").read().splitlines() if l.startswith("{")][-1])
print("exp2", {k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernels_ms"], "c2", d["extras"]["c2"]["ms_per_step"])
PY
timeout 300 python tools/e2e_profile.py c3 6 > gpurun_out/r2_e2e_profile_c3_6.txt 2>&1; head -36 gpurun_out/r2_e2e_profile_c3_6.txt

#!/bin/bash
# final validation of the round on one B200: GPU tests, smoke, bench lines, memcheck of the host-path entry points
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.txt
timeout 600 python bench.py > gpurun_out/bench_default.txt 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.txt 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 400 python bench.py --workload c2 --no-extras --no-cpu-baseline --e2e-videos 8 > gpurun_out/bench_c2.txt 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
for wl in c4_shard c4_trans; do
  timeout 600 python bench.py --workload $wl --steps 10 --e2e-videos 2 --no-extras --no-cpu-baseline > gpurun_out/bench_$wl.txt 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"
done
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitizer_memcheck.txt 2>&1; tail -2 gpurun_out/sanitizer_memcheck.txt
python - <<'PY'
import json
for f in ("bench_default", "bench_c2", "bench_c4_shard", "bench_c4_trans"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.txt").read().splitlines() if l.startswith("{")][-1])
        print(f, "%.4g" % d["value"], "%.4g ms" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"], "%.4g ms" % d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY

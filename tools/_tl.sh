cd "${GRAFT_REPO_ROOT:-.}"
free -g | head -2; grep -i -E "AnonHugePages|HugePages_Total|thp" /proc/meminfo | head; cat /sys/kernel/mm/transparent_hugepage/enabled
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "full_size or c_oracle_on_every or baseline" 2>&1 | tail -1
free -g | head -2
timeout 400 python bench.py --no-extras --no-cpu-baseline --steps 5 > gpurun_out/bench_quick.txt 2> gpurun_out/bench_quick.err; echo "quick rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_quick.txt").read().splitlines() if l.startswith("{")][-1])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "ms_each_step", "videos")})
PY
python tools/h2d_probe.py

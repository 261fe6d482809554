#!/bin/bash
# the default bench line and the reference arm, as the driver runs them
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_default.txt 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.txt 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_default.txt").read().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernel_ms"], {k: v for k, v in d["roofline"]["kernels_ms"].items() if k != "how"}, "frac", d["roofline"]["frac"], d["clocks"])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "ms_each_step", "videos")})
PY

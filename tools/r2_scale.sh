#!/bin/bash
# multi-GPU bench under torchrun: bash tools/r2_scale.sh N
N=$1
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.txt 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_n$N.txt").read().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["config"]["per_rank_ms_per_step"])
e = d["e2e"]; print("e2e", {k: e.get(k) for k in ("value", "ms_per_step", "ms_each_step", "videos", "records_equal_single_gpu_run", "records_checked_videos")})
PY
tail -3 gpurun_out/bench_n$N.err

cd "${GRAFT_REPO_ROOT:-.}"
A3D_BENCH_THREADS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > gpurun_out/bench_n2_t1.txt 2> gpurun_out/bench_n2_t1.err; echo rc=$?
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_n2_t1.txt").read().splitlines() if l.startswith("{")][-1])
e = d["e2e"]; print("e2e", {k: e.get(k) for k in ("value", "ms_per_step", "ms_each_step", "ms_each_step_per_rank")})
PY
tail -3 gpurun_out/bench_n2_t1.err

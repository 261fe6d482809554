#!/bin/bash
# Round-2 validation + profile capture on one B200 (run under gpurun from the repo root):
#   gpurun --timeout 3000 -- 'bash tools/r2_final.sh'
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
if [ -z "$SKIP_PYTEST" ]; then timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.txt; fi
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.txt
timeout 600 python bench.py > gpurun_out/bench_default.txt 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.txt 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 400 python bench.py --workload c2 --no-extras --no-cpu-baseline --e2e-videos 8 > gpurun_out/bench_c2.txt 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"
for wl in c4_shard c4_trans; do
  timeout 600 python bench.py --workload $wl --steps 10 --e2e-videos 2 --no-extras --no-cpu-baseline > gpurun_out/bench_$wl.txt 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"
done
for fr in 30 300; do timeout 300 python -m articulation3d_b200.tools.inference --output gpurun_out/inference_$fr --frames $fr --tracks 4 --save-obj 2>&1 | tail -1; done > gpurun_out/r2_inference_tool.txt
rm -rf gpurun_out/inference_30 gpurun_out/inference_300
tools/_build/pipes_bench > gpurun_out/r2_pipes_bench.txt 2>&1
M=smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_adu.sum,smsp__inst_executed_op_shared_atom.sum,sm__cycles_active.sum,gpu__time_duration.sum
timeout 300 ncu --clock-control none -k regex:k_project -s 1 -c 1 --metrics $M python tools/profile_pass.py --workload c3_mini --passes 2 > gpurun_out/r2_ncu_pipes_k_project.txt 2>&1
for spec in "c3_shard k_unproject" "c3_shard k_project" "c3_shard k_score_mma" "c3_shard k_finalize" "c2 k_unproject" "c2 k_project" "c2 k_score" "c2 k_finalize" "c3 k_project" "c3 k_score_mma" "c4_shard k_project" "c4_shard k_score_mma"; do
  set -- $spec
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f \
    -o gpurun_out/prof_$1_$2 python tools/profile_pass.py --workload $1 --passes 2 > gpurun_out/ncu_$1_$2.log 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_(unproject|project|score|finalize)" -c 200 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --workload c3_shard --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-videos 1 > gpurun_out/launches_bench.log 2>&1
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitizer_memcheck.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/sanitizer_racecheck.txt 2>&1
tail -2 gpurun_out/sanitizer_memcheck.txt gpurun_out/sanitizer_racecheck.txt
timeout 200 python tools/score_ab.py c3_shard ldg mma tma > gpurun_out/score_ab_c3_shard.txt 2>&1; tail -3 gpurun_out/score_ab_c3_shard.txt
python tools/refresh_profiles.py r2 > gpurun_out/refresh.log 2>&1; tail -1 gpurun_out/refresh.log | cut -c1-200
mkdir -p gpurun_out/profiles_new; cp profiles/r2_* profiles/traffic.json profiles/pipes.json gpurun_out/profiles_new/ 2>/dev/null
cp gpurun_out/r2_pipes_bench.txt gpurun_out/r2_ncu_pipes_k_project.txt gpurun_out/r2_inference_tool.txt gpurun_out/profiles_new/
find gpurun_out -name '*.ncu-rep' ! -name 'prof_c3_shard_k_project.ncu-rep' -delete; du -sh gpurun_out | tail -1

#!/bin/bash
# Round-end validation + profile capture on one B200 (run under gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/final_run.sh'
# then, back in the build container:  python tools/refresh_profiles.py
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
if [ -z "$SKIP_PYTEST" ]; then timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.txt; fi
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.txt
timeout 400 python bench.py > gpurun_out/bench_default.txt 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.txt 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
for wl in c2 c3_shard; do
  for k in k_unproject k_project k_score k_finalize; do
    A3D_SCORE_KERNEL=ldg timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f \
      -o gpurun_out/prof_${wl}_${k} python tools/profile_pass.py --workload $wl --passes 3 > gpurun_out/ncu_${wl}_${k}.log 2>&1
  done
done
A3D_PROJECT_KERNEL=exact timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_project -s 1 -c 1 -f \
  -o gpurun_out/prof_c3_shard_k_project_exact python tools/profile_pass.py --workload c3_shard --passes 3 > gpurun_out/ncu_c3_shard_k_project_exact.log 2>&1
A3D_SCORE_KERNEL=mma timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_score_mma -s 1 -c 1 -f \
  -o gpurun_out/prof_c3_shard_k_score_mma python tools/profile_pass.py --workload c3_shard --passes 3 > gpurun_out/ncu_c3_shard_k_score_mma.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/launches_bench_c2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-batched > gpurun_out/launches_bench_c2.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitizer_memcheck.txt 2>&1
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/sanitizer_racecheck.txt 2>&1
tail -2 gpurun_out/sanitizer_memcheck.txt gpurun_out/sanitizer_racecheck.txt
timeout 200 python tools/score_ab.py c3_shard ldg mma tma > gpurun_out/score_ab_c3_shard.txt 2>&1; tail -3 gpurun_out/score_ab_c3_shard.txt
for w in c2 c3_shard; do for m in 0 1 2; do timeout 200 python tools/project_ab.py $w $m 2>&1 | tail -2; done; done > gpurun_out/project_ab.txt 2>&1; tail -4 gpurun_out/project_ab.txt
A3D_LIB=$PWD/tools/_build/liba3d_stats.so timeout 200 python tools/filter_stats.py c2 c3_shard > gpurun_out/filter_stats.txt 2>&1; tail -3 gpurun_out/filter_stats.txt
# summaries are made here (the reports themselves are too big to travel back: 64 MiB limit on gpurun_out/)
python tools/refresh_profiles.py > gpurun_out/refresh.log 2>&1; mkdir -p gpurun_out/profiles_new; cp profiles/* gpurun_out/profiles_new/
find gpurun_out -name '*.ncu-rep' ! -name 'prof_c3_shard_k_project.ncu-rep' -delete; du -sh gpurun_out | tail -1

cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for wl in c3_mini c2; do
  for om in 0 1; do
    echo "== $wl per-lane list out_mode=$om"; A3D_LIB=$PWD/tools/_build/liba3d_nolist.so AB_ITERS=10 AB_OUT_MODE=$om timeout 300 python tools/project_ab.py $wl 2>&1 | tail -2
    echo "== $wl warp list out_mode=$om"; AB_ITERS=10 AB_OUT_MODE=$om timeout 300 python tools/project_ab.py $wl 2>&1 | tail -2
  done
done 2>&1 | tee gpurun_out/r2_ab12.txt
for m in 1 2; do echo "== c3_mini mode $m warp list"; AB_ITERS=10 AB_OUT_MODE=1 timeout 300 python tools/project_ab.py c3_mini $m 2>&1 | tail -2; echo "== per-lane"; A3D_LIB=$PWD/tools/_build/liba3d_nolist.so AB_ITERS=10 AB_OUT_MODE=1 timeout 300 python tools/project_ab.py c3_mini $m 2>&1 | tail -2; done 2>&1 | tee -a gpurun_out/r2_ab12.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt

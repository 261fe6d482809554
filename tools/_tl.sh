cd "${GRAFT_REPO_ROOT:-.}"
python tools/upload_probe.py 2>&1 | sed -n 2,2p
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tools.py tests/test_gpu_adapter.py -x -q -m gpu -k "optimize or tools or videos or adapter" 2>&1 | tail -3
TL_FINE=1 python tools/e2e_timeline.py c3 6 > gpurun_out/r2_e2e_timeline6.txt 2>&1
head -3 gpurun_out/r2_e2e_timeline6.txt
TL_QUIET=1 python tools/e2e_timeline.py c3 6 2>&1 | tail -3

cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
bash tools/ablate.sh run c3_mini 2>&1 | tee gpurun_out/r2_ablate_c3_mini.txt

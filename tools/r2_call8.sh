cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
python tools/h2d_probe.py 2>&1 | tee gpurun_out/r2_h2d_probe.txt
timeout 300 python tools/e2e_profile.py c3 6 > gpurun_out/r2_e2e_profile_c3_6.txt 2>&1; head -48 gpurun_out/r2_e2e_profile_c3_6.txt
for fr in 30 300; do timeout 300 python -m articulation3d_b200.tools.inference --output gpurun_out/inference_$fr --frames $fr --tracks 4 --save-obj 2>&1 | tail -1; done | tee gpurun_out/r2_inference_tool.txt
rm -rf gpurun_out/inference_30 gpurun_out/inference_300

#!/bin/bash
# one GPU call: pipe micro-benchmark, host profile of the public API, per-pipe instruction counts of k_project
tools/_build/pipes_bench > gpurun_out/r2_pipes_bench.txt 2>&1
cat gpurun_out/r2_pipes_bench.txt
timeout 300 python tools/e2e_profile.py c3 2 > gpurun_out/r2_e2e_profile_c3.txt 2>&1
head -70 gpurun_out/r2_e2e_profile_c3.txt
M=smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_fp16.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_adu.sum,smsp__inst_executed_op_shared_atom.sum,sm__cycles_active.sum,gpu__time_duration.sum
timeout 300 ncu --clock-control none -k regex:k_project -s 1 -c 1 --metrics $M python tools/profile_pass.py --workload c3_mini --passes 2 > gpurun_out/r2_ncu_pipes_k_project.txt 2>&1
tail -30 gpurun_out/r2_ncu_pipes_k_project.txt

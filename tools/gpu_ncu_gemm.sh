#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel --launch-skip 12 -c 3 -o gpurun_out/gemm_tc_pp -f python bench.py --workload pointpillars --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_gemm.log 2>&1
tail -5 gpurun_out/ncu_gemm.log; ls -la gpurun_out/*.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dense.py -q --timeout 300 > gpurun_out/r6_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r6_tests.log
# launch list of the 1-cloud forward (strong scaling at N = 8 runs this per GPU)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r6_launches_rl1.csv python bench.py --total-units 1 --steps 1 --warmup 3 --no-cpu > gpurun_out/r6_ncu_rl1.log 2>&1
python tools/launch_summary.py gpurun_out/r6_launches_rl1.csv --all 2>&1 | head -75
# full captures: the 8 LFA launches of one 8-cloud step, and the PointPillars dense kernels
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lfa -c 8 -o gpurun_out/r6_lfa python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r6_ncu_lfa.log 2>&1
echo "ncu lfa rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc -c 20 -o gpurun_out/r6_gemm python bench.py --workload pointpillars --steps 1 --warmup 3 --no-cpu > gpurun_out/r6_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# Final evidence run of the round (1 GPU): GPU tests, the three bench lines, the reference arm, launch lists,
# the in-situ GEMM timeline.  Results land in gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/final_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
for w in randlanet pointpillars kpconv; do
  timeout 900 python bench.py --workload $w > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  echo "bench $w rc=$?"; cut -c1-260 gpurun_out/r02_bench_$w.json
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-200 gpurun_out/r02_bench_reference.json
timeout 600 python bench.py --workload pointpillars --shape waymo --total-units 4 --no-cpu > gpurun_out/r02_bench_pointpillars_waymo4.json 2> gpurun_out/r02_bench_pointpillars_waymo4.err; cut -c1-200 gpurun_out/r02_bench_pointpillars_waymo4.json
timeout 600 python bench.py --total-units 1 --no-cpu > gpurun_out/r02_bench_randlanet_1cloud.json 2> /dev/null; cut -c1-200 gpurun_out/r02_bench_randlanet_1cloud.json
for w in randlanet pointpillars kpconv; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
done
O3DML_DEBUG_TIMING=1 python open3d-ml_b200/build.py --force > /dev/null 2>&1
timeout 300 python tools/debug_timeline_insitu.py > gpurun_out/r02_gemm_tc_timeline.txt 2>&1; tail -16 gpurun_out/r02_gemm_tc_timeline.txt

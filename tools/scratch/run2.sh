#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -q --timeout 120 -x > gpurun_out/r2_dense.log 2>&1
echo "dense rc=$?"; tail -5 gpurun_out/r2_dense.log
timeout 900 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/r2_all.log 2>&1
echo "all rc=$?"; tail -25 gpurun_out/r2_all.log
for w in pointpillars randlanet kpconv; do
  timeout 400 python bench.py --workload $w --no-cpu > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err
  echo "bench $w rc=$?"; cut -c1-330 gpurun_out/r2_bench_$w.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_pp.csv python bench.py --workload pointpillars --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_ncu_pp.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_pp.csv --all 2>&1 | tail -50
# in-situ timeline with the debug build
O3DML_DEBUG_TIMING=1 python open3d-ml_b200/build.py --force > /dev/null 2>&1
timeout 300 python tools/debug_timeline_insitu.py 2>&1 | tee gpurun_out/r2_timeline.txt | tail -20

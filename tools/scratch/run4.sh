#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_models.py -q --timeout 180 > gpurun_out/r4_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r4_tests.log
timeout 1500 python -m pytest tests/test_gpu_reference_boundary.py -q --timeout 900 > gpurun_out/r4_boundary.log 2>&1
echo "boundary rc=$?"; tail -30 gpurun_out/r4_boundary.log
for w in pointpillars kpconv; do
  timeout 600 python bench.py --workload $w --no-cpu > gpurun_out/r4_bench_$w.json 2> gpurun_out/r4_bench_$w.err
  echo "bench $w rc=$?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r4_bench_$w.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
except Exception as e: print('ERR', e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r4_launches_knn.csv python tools/knn_profile.py > gpurun_out/r4_knn.log 2>&1
tail -2 gpurun_out/r4_knn.log
python tools/launch_summary.py gpurun_out/r4_launches_knn.csv 2>&1 | head -30
O3DML_DEBUG_TIMING=1 python open3d-ml_b200/build.py --force > /dev/null 2>&1
timeout 300 python tools/debug_timeline_insitu.py 2>&1 | tee gpurun_out/r4_timeline.txt | tail -20

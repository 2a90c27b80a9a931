#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dense.py -q --timeout 120 -x > gpurun_out/r10_ops.log 2>&1
echo "ops rc=$?"; tail -6 gpurun_out/r10_ops.log
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_lfa_tc.py -q --timeout 900 > gpurun_out/r10_models.log 2>&1
echo "models rc=$?"; tail -4 gpurun_out/r10_models.log
run() {
timeout 600 python bench.py $2 --no-cpu > gpurun_out/r10_$1.json 2> gpurun_out/r10_$1.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r10_$1.json')); print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['entry'], {k:v['value'] for k,v in d['e2e']['other_entries'].items()}, d.get('knn_pyramid_ms'), d.get('kpconv_batch_build_ms'))
except Exception as e: print('ERR', e); print(open('gpurun_out/r10_$1.err').read()[-1500:])
PY
}
run pp "--workload pointpillars"
run rl ""
run rl1 "--total-units 1"
run kp "--workload kpconv"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r10_launches_knn.csv python tools/knn_profile.py > gpurun_out/r10_knn.log 2>&1
tail -1 gpurun_out/r10_knn.log; python tools/launch_summary.py gpurun_out/r10_launches_knn.csv 2>&1 | head -8

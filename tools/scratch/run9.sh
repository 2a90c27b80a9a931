#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q --timeout 120 -x > gpurun_out/r9_ops.log 2>&1
echo "ops rc=$?"; tail -6 gpurun_out/r9_ops.log
timeout 900 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r9_all.log 2>&1
echo "all rc=$?"; tail -6 gpurun_out/r9_all.log
for w in pointpillars randlanet; do
timeout 600 python bench.py --workload $w --no-cpu > gpurun_out/r9_$w.json 2> gpurun_out/r9_$w.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r9_$w.json')); print('$w', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('knn_pyramid_ms'))
except Exception as e: print('ERR', e); print(open('gpurun_out/r9_$w.err').read()[-1500:])
PY
done

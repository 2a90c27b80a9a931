#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lfa_tc.py tests/test_gpu_models.py -q --timeout 300 > gpurun_out/r8_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r8_tests.log
timeout 600 python bench.py --no-cpu > gpurun_out/r8_rl.json 2> gpurun_out/r8_rl.err
python - <<PY
import json
d=json.load(open('gpurun_out/r8_rl.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value']); r=d['roofline']; print(r['frac'], r['share_of_step'], {k:v['avg_us'] for k,v in r['per_kernel'].items()})
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_reference_boundary.py tests/test_gpu_lfa_tc.py tests/test_gpu_models.py -q --timeout 900 > gpurun_out/r5_tests.log 2>&1
echo "tests rc=$?"; tail -30 gpurun_out/r5_tests.log
run() { # name, extra env, args
  env $2 timeout 600 python bench.py $3 --no-cpu > gpurun_out/r5_$1.json 2> gpurun_out/r5_$1.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r5_$1.json')); print('$1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['scaling'], d.get('knn_pyramid_ms'))
except Exception as e: print('$1 ERR', e); print(open('gpurun_out/r5_$1.err').read()[-1500:])
PY
}
run rl_1cloud "A=1" "--total-units 1"
run rl_2cloud "A=1" "--total-units 2"
run kp_k32 "O3DML_GEMM_TC_MIN_K=32" "--workload kpconv"
run kp_k64 "O3DML_GEMM_TC_MIN_K=64" "--workload kpconv"
run rl_k32 "O3DML_GEMM_TC_MIN_K=32" ""
run rl_k64 "O3DML_GEMM_TC_MIN_K=64" ""
run pp_waymo "A=1" "--workload pointpillars --shape waymo --total-units 4"

#!/bin/bash
# one GPU iteration: the LFA / model parity tests, then the three bench workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lfa_tc.py tests/test_gpu_models.py -q --tb=short -p no:cacheprovider -x > gpurun_out/quick.log 2>&1; echo "exit $?" >> gpurun_out/quick.log; tail -8 gpurun_out/quick.log
bash tools/gpu_bench3.sh

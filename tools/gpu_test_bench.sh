#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_models.py -q --tb=short -p no:cacheprovider > gpurun_out/quick.log 2>&1; echo "exit $?" >> gpurun_out/quick.log; tail -6 gpurun_out/quick.log
bash tools/gpu_bench_light.sh

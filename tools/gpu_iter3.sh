#!/bin/bash
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_models.py -q --tb=short -p no:cacheprovider -x > gpurun_out/quick.log 2>&1; echo "exit $?" >> gpurun_out/quick.log; tail -8 gpurun_out/quick.log
timeout 300 python tools/pp_layer_times.py > gpurun_out/pp_layers_$R.txt 2>&1; cat gpurun_out/pp_layers_$R.txt | tail -24
bash tools/gpu_bench3.sh

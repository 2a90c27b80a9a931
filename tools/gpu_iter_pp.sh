#!/bin/bash
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 600 python -m pytest tests/test_gpu_dense.py -q --tb=short -p no:cacheprovider -x -k "linear or conv" > gpurun_out/quick.log 2>&1; echo "exit $?" >> gpurun_out/quick.log; tail -4 gpurun_out/quick.log
timeout 300 python tools/pp_layer_times.py > gpurun_out/pp_layers_$R.txt 2>&1; cat gpurun_out/pp_layers_$R.txt | tail -24
timeout 600 python bench.py --workload kpconv --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_kpconv_$R.json 2> gpurun_out/bench_kpconv.err; head -c 200 gpurun_out/bench_kpconv_$R.json

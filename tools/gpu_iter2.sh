#!/bin/bash
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_models.py -q --tb=short -p no:cacheprovider -x > gpurun_out/quick.log 2>&1; echo "exit $?" >> gpurun_out/quick.log; tail -8 gpurun_out/quick.log
timeout 300 python tools/pp_layer_times.py > gpurun_out/pp_layers_$R.txt 2>&1; cat gpurun_out/pp_layers_$R.txt | tail -30
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_randlanet_$R.json 2> gpurun_out/bench_randlanet.err; tail -c 1600 gpurun_out/bench_randlanet_$R.json; tail -3 gpurun_out/bench_randlanet.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_randlanet_$R.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch_randlanet.log 2>&1

#!/bin/bash
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_models.py -q --tb=short -p no:cacheprovider -x -k "kp or gather" > gpurun_out/quick.log 2>&1; echo "exit $?" >> gpurun_out/quick.log; tail -8 gpurun_out/quick.log
timeout 900 python bench.py --workload kpconv --steps 10 --warmup 3 > gpurun_out/bench_kpconv_$R.json 2> gpurun_out/bench_kpconv.err; tail -c 1200 gpurun_out/bench_kpconv_$R.json; tail -8 gpurun_out/bench_kpconv.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_kpconv_$R.csv python bench.py --workload kpconv --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch_kp.log 2>&1

#!/bin/bash
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_randlanet_$R.json 2> gpurun_out/bench_randlanet.err; tail -c 2300 gpurun_out/bench_randlanet_$R.json; tail -3 gpurun_out/bench_randlanet.err
timeout 600 python bench.py --workload pointpillars --steps 20 --warmup 5 > gpurun_out/bench_pointpillars_$R.json 2> gpurun_out/bench_pointpillars.err; tail -c 2300 gpurun_out/bench_pointpillars_$R.json; tail -3 gpurun_out/bench_pointpillars.err
timeout 900 python bench.py --workload kpconv --steps 10 --warmup 3 > gpurun_out/bench_kpconv_$R.json 2> gpurun_out/bench_kpconv.err; tail -c 2300 gpurun_out/bench_kpconv_$R.json; tail -8 gpurun_out/bench_kpconv.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_kpconv_$R.csv python bench.py --workload kpconv --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch_kp.log 2>&1

#!/bin/bash
# bench lines + launch lists (no full ncu capture, no pytest)
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_randlanet_$R.json 2> gpurun_out/bench_randlanet.err; tail -c 2600 gpurun_out/bench_randlanet_$R.json; tail -5 gpurun_out/bench_randlanet.err
timeout 600 python bench.py --workload pointpillars --steps 20 --warmup 5 > gpurun_out/bench_pointpillars_$R.json 2> gpurun_out/bench_pointpillars.err; tail -c 1500 gpurun_out/bench_pointpillars_$R.json; tail -5 gpurun_out/bench_pointpillars.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_randlanet_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_pointpillars_$R.csv python bench.py --workload pointpillars --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_pp.log 2>&1

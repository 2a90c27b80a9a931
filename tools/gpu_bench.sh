#!/bin/bash
# Runs on the GPU box: parity suite, smoke, bench lines, ncu launch list + one full capture.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_randlanet_$R.json 2> gpurun_out/bench_randlanet.err; tail -c 3000 gpurun_out/bench_randlanet_$R.json; tail -5 gpurun_out/bench_randlanet.err
timeout 600 python bench.py --workload pointpillars --steps 20 --warmup 5 > gpurun_out/bench_pointpillars_$R.json 2> gpurun_out/bench_pointpillars.err; tail -c 2000 gpurun_out/bench_pointpillars_$R.json; tail -5 gpurun_out/bench_pointpillars.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$R.json 2>&1; tail -c 1500 gpurun_out/bench_reference_$R.json
if [ -z "$NO_NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_randlanet_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_pointpillars_$R.csv python bench.py --workload pointpillars --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_pp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lfa_pool -c 8 -o gpurun_out/lfa_$R -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
fi

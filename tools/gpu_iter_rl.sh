#!/bin/bash
# LFA / RandLA-Net iteration: parity tests, then the default bench
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 600 python -m pytest tests/test_gpu_lfa_tc.py tests/test_gpu_models.py -q --tb=short -p no:cacheprovider -x -k "lfa or randla or pipelined" > gpurun_out/quick.log 2>&1; echo "exit $?" >> gpurun_out/quick.log; tail -12 gpurun_out/quick.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_randlanet_$R.json 2> gpurun_out/bench_randlanet.err; tail -c 1500 gpurun_out/bench_randlanet_$R.json; tail -3 gpurun_out/bench_randlanet.err

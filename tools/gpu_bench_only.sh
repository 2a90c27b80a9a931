#!/bin/bash
mkdir -p gpurun_out
R=${ROUND:-r01}
for W in randlanet pointpillars kpconv; do
timeout 600 python bench.py --workload $W --steps 20 --warmup 5 > gpurun_out/bench_${W}_$R.json 2> gpurun_out/bench_$W.err; head -c 160 gpurun_out/bench_${W}_$R.json; echo; tail -2 gpurun_out/bench_$W.err
done

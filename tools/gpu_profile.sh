#!/bin/bash
# profiles for the round: per-launch list of one timed step per workload + one ncu --set full
# capture of the LFA kernels (8 launches = one RandLA-Net step)
mkdir -p gpurun_out
R=${ROUND:-r01}
for W in randlanet pointpillars kpconv; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${W}_$R.csv python bench.py --workload $W --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch_$W.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lfa -c 8 -o gpurun_out/lfa_$R -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*_$R.csv

#!/bin/bash
# Final evidence run of the round (1 GPU): GPU tests, smoke, the three bench lines, the reference arm, the one-cloud
# line, launch lists, one ncu --set full capture of the LFA kernels.  Results land in gpurun_out/ (copied into
# profiles/ afterwards; tools/ncu_report.py turns the capture into the markdown table + traffic json).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/final_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
for w in randlanet pointpillars kpconv; do
  timeout 600 python bench.py --workload $w > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  echo "bench $w rc=$?"; cut -c1-260 gpurun_out/r02_bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-200 gpurun_out/r02_bench_reference.json
timeout 300 python bench.py --total-units 1 --no-cpu > gpurun_out/r02_bench_randlanet_1cloud.json 2> /dev/null; cut -c1-200 gpurun_out/r02_bench_randlanet_1cloud.json
for w in randlanet pointpillars kpconv; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lfa -c 8 -o gpurun_out/r02_lfa_final -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/r02_lfa_final.ncu-rep

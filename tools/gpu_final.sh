#!/bin/bash
# round-end evidence: whole GPU suite, smoke(), three bench lines, per-launch lists, one ncu --set full
# capture of the LFA kernels and one of the top gemm_tc launches
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
for W in randlanet pointpillars kpconv; do
timeout 900 python bench.py --workload $W --steps 20 --warmup 5 > gpurun_out/bench_${W}_$R.json 2> gpurun_out/bench_$W.err; tail -c 400 gpurun_out/bench_${W}_$R.json; tail -2 gpurun_out/bench_$W.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${W}_$R.csv python bench.py --workload $W --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch_$W.log 2>&1
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$R.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference_$R.json
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lfa -c 8 -o gpurun_out/lfa_$R -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel --launch-skip 11 -c 2 -o gpurun_out/gemm_tc_pp_$R -f python bench.py --workload pointpillars --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_full.log gpurun_out/ncu_gemm.log; ls -la gpurun_out/*_$R.*

#!/bin/bash
# GPU run 1: validate the TF32/TMA gemm_tc kernel, then measure.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dense.py -q --timeout 120 -x > gpurun_out/r1_dense.log 2>&1
echo "dense rc=$?" | tee -a gpurun_out/r1_dense.log
tail -25 gpurun_out/r1_dense.log
timeout 900 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/r1_all.log 2>&1
echo "all rc=$?" | tee -a gpurun_out/r1_all.log
tail -15 gpurun_out/r1_all.log
timeout 300 python tools/pp_layer_times.py > gpurun_out/r1_pp_layers.txt 2>&1; tail -30 gpurun_out/r1_pp_layers.txt
for w in pointpillars randlanet kpconv; do
  timeout 400 python bench.py --workload $w --no-cpu > gpurun_out/r1_bench_$w.json 2> gpurun_out/r1_bench_$w.err
  echo "bench $w rc=$?"; cat gpurun_out/r1_bench_$w.json | cut -c1-600
done

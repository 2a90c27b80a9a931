#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/scratch/repro_b2.py > gpurun_out/r12_repro.log 2>&1; echo "repro rc=$?"; tail -15 gpurun_out/r12_repro.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/scratch/repro_b2.py > gpurun_out/r12_sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -v "^$" gpurun_out/r12_sanitizer.log | grep -A25 "Invalid\|ERROR SUMMARY\|=========" | head -70 | cut -c1-220
timeout 600 python -m pytest tests/test_gpu_models.py -q --timeout 300 -k "tail or five_level or forward_points" > gpurun_out/r12_tail.log 2>&1; echo "tail tests rc=$?"; tail -12 gpurun_out/r12_tail.log | cut -c1-250

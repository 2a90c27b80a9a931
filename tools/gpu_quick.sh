#!/bin/bash
# quick GPU check of a subset: bash tools/gpu_quick.sh <pytest args>
mkdir -p gpurun_out
timeout 600 python -m pytest "$@" -q --tb=short -p no:cacheprovider > gpurun_out/quick.log 2>&1
echo "exit $?" >> gpurun_out/quick.log
tail -40 gpurun_out/quick.log

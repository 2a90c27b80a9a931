#!/bin/bash
# 1 -> 8 GPU scaling lines of the default workload on one 8-GPU box (strong, the default; weak at N=8).
mkdir -p gpurun_out
run() { # n, extra args, tag
  n=$1; extra=$2; tag=$3
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 $extra 2>gpurun_out/scale_${tag}.err | tail -1 > gpurun_out/r02_scale_${tag}.json
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 30 --warmup 5 $extra 2>gpurun_out/scale_${tag}.err | grep '^{' | tail -1 > gpurun_out/r02_scale_${tag}.json
  fi
  echo "$tag rc=$? $(head -c 200 gpurun_out/r02_scale_${tag}.json)"
}
run 1 "--no-cpu" n1_strong
run 2 "" n2_strong
run 4 "" n4_strong
run 8 "" n8_strong
run 8 "--units 8" n8_weak

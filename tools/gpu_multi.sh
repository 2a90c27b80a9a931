#!/bin/bash
# multi-GPU bench: N ranks on one box via torch.distributed.run
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_randlanet_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1500 gpurun_out/bench_randlanet_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 1 --impl reference > gpurun_out/bench_reference_n$N.json 2>> gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_reference_n$N.json

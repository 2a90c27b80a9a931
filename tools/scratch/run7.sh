#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -5
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dense.py tests/test_gpu_lfa_tc.py tests/test_gpu_models.py -q --timeout 300 > gpurun_out/r7_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r7_tests.log
show() { python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$1.json')); print('$1', 'value', d['value'], 'ms', d['ms_per_step'], d['scaling'], 'n', d['n_gpus'], 'e2e', d['e2e']['value'], d['e2e']['entry'], {k:v['value'] for k,v in d['e2e']['other_entries'].items()}, d['e2e']['collective'][:20])
except Exception as e: print('$1 ERR', e); print(open('gpurun_out/$1.err').read()[-2500:])
PY
}
timeout 600 python bench.py --no-cpu > gpurun_out/r7_n1.json 2> gpurun_out/r7_n1.err; show r7_n1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu > gpurun_out/r7_n2.json 2> gpurun_out/r7_n2.err; show r7_n2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --units 8 --no-cpu > gpurun_out/r7_n2_weak.json 2> gpurun_out/r7_n2_weak.err; show r7_n2_weak
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r7_n2_ref.json 2> gpurun_out/r7_n2_ref.err; cut -c1-300 gpurun_out/r7_n2_ref.json
timeout 600 python bench.py --workload pointpillars --no-cpu > gpurun_out/r7_pp.json 2> gpurun_out/r7_pp.err; show r7_pp
timeout 600 python bench.py --workload kpconv --no-cpu > gpurun_out/r7_kp.json 2> gpurun_out/r7_kp.err; show r7_kp

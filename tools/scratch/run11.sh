#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
show() { python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$1.json')); print('$1', 'value', d['value'], 'ms', d['ms_per_step'], d['scaling'], 'n', d['n_gpus'], 'e2e', d['e2e']['value'], d['e2e']['entry'], {k:v['value'] for k,v in d['e2e']['other_entries'].items()}, 'numa', d['config'].get('numa_node_rank0'))
except Exception as e: print('$1 ERR', e); print(open('gpurun_out/$1.err').read()[-2500:])
PY
}
P=29600
for spec in "r02_scale_n8_strong:8:" "r02_scale_n4_strong:4:" "r02_scale_n8_weak:8:--units 8" "r02_bench_pointpillars_waymo32_n8:8:--workload pointpillars --shape waymo --total-units 32"; do
  name=${spec%%:*}; rest=${spec#*:}; n=${rest%%:*}; extra=${rest#*:}
  P=$((P+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n $extra --no-cpu > gpurun_out/$name.json 2> gpurun_out/$name.err
  show $name
done

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29901 bench.py --gpus 4 --no-cpu > gpurun_out/r02_scale_n4_strong.json 2> gpurun_out/r02_scale_n4_strong.err
echo "n4 rc=$?"; cut -c1-200 gpurun_out/r02_scale_n4_strong.json; grep -c "illegal memory" gpurun_out/r02_scale_n4_strong.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_scale_n4_strong.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['entry'], {k:v['value'] for k,v in d['e2e']['other_entries'].items()}, d['e2e']['d2h_bytes_per_step'])
PY

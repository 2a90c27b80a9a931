#!/bin/bash
# 2-GPU sanity of the sharded bench: strong (8 clouds over 2 ranks) and the 2-clouds-per-rank weak configuration that
# faulted in the first multi-GPU session of the round (profiles/README.md)
mkdir -p gpurun_out
: > gpurun_out/n2_check.log
for a in "" "--units 2"; do
  timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 $a 2>gpurun_out/n2_check.err | grep "^{" | tail -1 >> gpurun_out/n2_check.log
  echo "rc=$?"
done
python - <<'PY'
import json
for l in open("gpurun_out/n2_check.log"):
    try:
        d = json.loads(l)
        print(d["scaling"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["entry"])
    except Exception as e:
        print("bad line", l[:300])
PY
tail -5 gpurun_out/n2_check.err

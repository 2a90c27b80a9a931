#!/bin/bash
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print(sys.argv[1], d["value"], d["ms_per_step"], {k.replace("lfa_pool",""):v["avg_us"] for k,v in pk.items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() { tag=$1; flags=$2; O3DML_NVCC_EXTRA="$flags" python open3d-ml_b200/build.py --force > /dev/null 2>gpurun_out/ab_$tag.build || echo build failed; 
  O3DML_NVCC_EXTRA="$flags" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu 2>gpurun_out/ab_$tag.err | tail -1 > gpurun_out/ab_$tag.json; summ gpurun_out/ab_$tag.json; }
run base ""
run eager "-DLTC_EAGER_INDEX"
run expf "-DLTC_EXPF"
run l16c4 "-DL16C_CTAS=4"
run l16c6 "-DL16C_CTAS=6"
run base2 ""

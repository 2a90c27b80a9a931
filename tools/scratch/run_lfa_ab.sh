#!/bin/bash
# A/B of the LFA gather pipelining / LocSE-under-MMA variants (development run)
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_lfa_tc.py tests/test_gpu_models.py -k "lfa or randlanet" -q -x 2>&1 | tail -8 > gpurun_out/ab_tests.log
cat gpurun_out/ab_tests.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print(sys.argv[1], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], {k.replace("lfa_pool",""):v["avg_us"] for k,v in pk.items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu 2>gpurun_out/ab_$tag.err | tail -1 > gpurun_out/ab_$tag.json; summ gpurun_out/ab_$tag.json; }
run default X=1
run noshadow O3DML_LFA_SHADOW=0
run pf0 O3DML_LFA16_PF=0
run pf2 O3DML_LFA16_PF=2
run default2 X=1

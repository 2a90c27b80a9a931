#!/bin/bash
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    pk=d["roofline"].get("per_kernel",{})
    print(sys.argv[1], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], {k.replace("lfa_pool",""):v["avg_us"] for k,v in pk.items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() { tag=$1; flags=$2; shift; shift; O3DML_NVCC_EXTRA="$flags" python open3d-ml_b200/build.py --force > /dev/null 2>gpurun_out/abc_$tag.build || echo build failed;
  O3DML_NVCC_EXTRA="$flags" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu "$@" 2>gpurun_out/abc_$tag.err | tail -1 > gpurun_out/abc_$tag.json; summ gpurun_out/abc_$tag.json; }
python open3d-ml_b200/build.py --force > /dev/null
timeout 600 python -m pytest tests/test_gpu_lfa_tc.py tests/test_gpu_models.py tests/test_gpu_dense.py -q -x 2>&1 | tail -5
run base ""
run pp "" --workload pointpillars
run kp "" --workload kpconv
run rawall "-DLTC_RAW_INDEX_ALL"

#!/bin/bash
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
python open3d-ml_b200/build.py > /dev/null
O3DML_GEMM_LITE=100000 O3DML_GEMM_LITE_CONV=1 timeout 400 python -m pytest tests/test_gpu_models.py tests/test_gpu_dense.py -q -x -k "pointpillars or conv or kpfcnn or linear_plain" 2>&1 | tail -4
i=0
for v in "O3DML_GEMM_LITE=16" "O3DML_GEMM_LITE=64" "O3DML_GEMM_LITE=100000" "O3DML_GEMM_LITE=100000 O3DML_GEMM_LITE_CONV=1"; do
  i=$((i+1))
  for wl in randlanet pointpillars kpconv; do
    env $v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --workload $wl 2>gpurun_out/l2_${i}_$wl.err | tail -1 > gpurun_out/l2_${i}_$wl.json
    echo -n "$v "; summ gpurun_out/l2_${i}_$wl.json
  done
done

#!/bin/bash
# parity tests of the dense / LFA / model paths, then the three bench workloads under a list of environment variants
# (development A/B in one box session).  usage: bash tools/gpu_ab_check.sh "VAR=a" "VAR=b" ...
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
timeout 700 python -m pytest tests/test_gpu_dense.py tests/test_gpu_lfa_tc.py tests/test_gpu_models.py -q -x 2>&1 | tail -5
i=0
for v in "$@"; do
  i=$((i+1))
  for wl in randlanet pointpillars kpconv; do
    env $v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --workload $wl 2>gpurun_out/abc_${i}_$wl.err | tail -1 > gpurun_out/abc_${i}_$wl.json
    echo -n "$v "; summ gpurun_out/abc_${i}_$wl.json
  done
done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/r3_all.log 2>&1
echo "all rc=$?"; tail -25 gpurun_out/r3_all.log
for w in randlanet pointpillars kpconv; do
  timeout 600 python bench.py --workload $w > gpurun_out/r3_bench_$w.json 2> gpurun_out/r3_bench_$w.err
  echo "bench $w rc=$?"; tail -3 gpurun_out/r3_bench_$w.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r3_bench_$w.json'))
    print({k:d[k] for k in d if k not in ('config','clocks','roofline','e2e')})
    print('e2e', json.dumps(d['e2e'])[:900])
    r=d['roofline']; print('roofline', {k:r[k] for k in r if k in ('achieved','frac','share_of_step','fp32_tflops')})
except Exception as e: print('ERR', e)
PY
done
O3DML_DEBUG_TIMING=1 python open3d-ml_b200/build.py --force > /dev/null 2>&1
timeout 300 python tools/debug_timeline_insitu.py 2>&1 | tee gpurun_out/r3_timeline.txt | tail -20

#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/tc_diag.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from test_gpu_tc import run
for variant in (0, 1):
    for (n, k) in [(16, 16), (64, 64), (128, 128), (256, 128)]:
        for terms in (1, 3):
            try:
                err, d, ref = run(n, k, terms | (variant << 8))
                print("variant", variant, "n", n, "k", k, "terms", terms, "rel_err %.3e" % err, "nan" if torch.isnan(d).any() else "")
            except Exception as e:
                print("variant", variant, n, k, terms, "EXC", e)
PY
cat gpurun_out/tc_diag.log
timeout 300 python -m pytest tests/test_gpu_tc.py -q --tb=short -p no:cacheprovider 2>&1 | tail -15

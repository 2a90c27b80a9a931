#!/bin/bash
# repeat the RandLA parity tests several times (timing-dependent bugs)
for i in 1 2 3; do
timeout 300 python tools/debug_lfa.py 2>&1 | grep -E "DBG|d 256"
done
timeout 900 python -m pytest tests/test_gpu_lfa_tc.py tests/test_gpu_models.py -q --tb=short -p no:cacheprovider -k "lfa_tc or randlanet" 2>&1 | tail -8

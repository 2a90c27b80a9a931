#!/bin/bash
timeout 300 python tools/debug_rate.py 2>&1 | tail -32
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_models.py -q --tb=short -p no:cacheprovider 2>&1 | tail -6

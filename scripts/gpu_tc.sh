#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=line > gpurun_out/pytest_tc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tc.log
grep -E "passed|failed" gpurun_out/pytest_tc.log | tail -3
grep -E "^FAILED|^/root|^E " gpurun_out/pytest_tc.log | sed 's/ - .*//' | head -60

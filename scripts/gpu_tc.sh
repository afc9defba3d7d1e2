#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short -x -k "${K:-conv or cifar}" > gpurun_out/pytest_tc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log

#!/bin/bash
# full GPU check: pytest -m gpu, smoke(), default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/r2_smoke.log | cut -c1-600
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench exit $?"; tail -5 gpurun_out/r2_bench.err; cut -c1-1500 gpurun_out/r2_bench.json

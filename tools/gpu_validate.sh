#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log)
(timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu.log | tail -8)
(timeout 1500 python bench.py > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; echo "bench (default flags) exit $?"; cat gpurun_out/bench_large.json; tail -3 gpurun_out/bench_large.err)

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|Error" gpurun_out/pytest_gpu.log | tail -8)
(timeout 300 python tools/profile_step.py --streams 32 --tokens 24 > gpurun_out/step32.log 2>&1; echo "PDL on: $(tail -1 gpurun_out/step32.log)")
(WLB200_PDL=0 timeout 300 python tools/profile_step.py --streams 32 --tokens 24 > gpurun_out/step32_nopdl.log 2>&1; echo "PDL off: $(tail -1 gpurun_out/step32_nopdl.log)")
(timeout 300 python tools/profile_step.py --streams 8 --tokens 24 > gpurun_out/step8.log 2>&1; echo "8 streams: $(tail -1 gpurun_out/step8.log)")
(timeout 300 python tools/kbench.py > gpurun_out/kbench.txt 2>&1; tail -30 gpurun_out/kbench.txt)
(timeout 1500 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; echo "bench large exit $?"; cat gpurun_out/bench_large.json; tail -5 gpurun_out/bench_large.err)

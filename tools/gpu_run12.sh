#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu.log | tail -12)
for ns in 0 1 2 3 4 6; do
(WLB200_XA_NSPLIT=$ns timeout 300 python tools/profile_step.py --streams 32 --tokens 24 > gpurun_out/step32_ns$ns.log 2>&1; echo "nsplit=$ns: $(tail -1 gpurun_out/step32_ns$ns.log | sed 's/.*mel ms/mel ms/')")
done
(timeout 300 python tools/profile_step.py --streams 8 --tokens 24 > gpurun_out/step8.log 2>&1; echo "8 streams: $(tail -1 gpurun_out/step8.log | sed 's/.*mel ms/mel ms/')")
(timeout 1500 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; echo "bench large exit $?"; cat gpurun_out/bench_large.json; tail -5 gpurun_out/bench_large.err)

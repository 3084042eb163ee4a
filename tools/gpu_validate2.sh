#!/bin/bash
# Round-2 validation: smoke, GPU suite, default bench, then the per-rank operating points of the scaling series and config 4.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log)
(timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|Error|assert|decode session|step rounds" gpurun_out/pytest_gpu.log | tail -14)
(timeout 600 python bench.py > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; rc=$?; echo "bench (default flags) exit $rc"; cat gpurun_out/bench_large.json; tail -3 gpurun_out/bench_large.err
 if [ $rc -ne 0 ]; then timeout 600 python bench.py --no-streaming > gpurun_out/bench_large_ns.json 2> gpurun_out/bench_large_ns.err; echo "bench --no-streaming exit $?"; cat gpurun_out/bench_large_ns.json; tail -3 gpurun_out/bench_large_ns.err; fi)
for s in 4 8; do
(timeout 300 python bench.py --streams $s --no-cpu-baseline > gpurun_out/bench_s$s.json 2> gpurun_out/bench_s$s.err; echo "bench streams $s exit $?"; cat gpurun_out/bench_s$s.json; tail -2 gpurun_out/bench_s$s.err)
done
(timeout 400 python bench.py --word-timestamps --no-cpu-baseline > gpurun_out/bench_wt.json 2> gpurun_out/bench_wt.err; echo "bench word-timestamps exit $?"; cat gpurun_out/bench_wt.json; tail -2 gpurun_out/bench_wt.err)

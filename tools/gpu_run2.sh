#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 300 python tools/debug_gen.py > gpurun_out/debug_gen.log 2>&1; echo "debug_gen exit $?")
(timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log)
(timeout 300 python tools/kbench.py > gpurun_out/kbench.txt 2>&1; echo "kbench exit $?"; cat gpurun_out/kbench.txt)
(WLB200_BN=256 timeout 300 python tools/kbench.py > gpurun_out/kbench_bn256.txt 2>&1; head -8 gpurun_out/kbench_bn256.txt)
(timeout 600 python bench.py --model small.en --streams 4 --steps 2 --warmup 1 --beam 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench small exit $?"; cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err)
(timeout 1200 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; echo "bench large exit $?"; cat gpurun_out/bench_large.json; tail -5 gpurun_out/bench_large.err)

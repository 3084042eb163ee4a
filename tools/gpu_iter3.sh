#!/bin/bash
# flash-attention (two alternating softmax groups) validation + encoder timing + prefetch on the split-K path + default bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -s -k "encoder or full_size or small_en or mel" > gpurun_out/pytest_enc.log 2>&1; echo "pytest encoder exit $?"; grep -E "passed|failed|Error|assert|max err|rel rms" gpurun_out/pytest_enc.log | tail -12)
(timeout 200 python tools/enc_time.py --streams 32 --reps 3 2>&1 | tee gpurun_out/enc_time.txt | tail -4)
(timeout 200 python tools/sweep_prefetch.py --streams 32 --cgemm 0 --values 0,8,16 --tokens 40 --reps 2 2>&1 | tee gpurun_out/sweep_prefetch2.txt | tail -4)
(timeout 600 python bench.py > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; echo "bench (default flags) exit $?"; cat gpurun_out/bench_large.json; tail -3 gpurun_out/bench_large.err)

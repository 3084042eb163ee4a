#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu.log | tail -12)
(timeout 1500 python bench.py --steps 3 --warmup 1 > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; echo "bench large exit $?"; cat gpurun_out/bench_large.json; tail -5 gpurun_out/bench_large.err)
(timeout 300 python tools/profile_step.py --streams 32 --tokens 24 > gpurun_out/step32.log 2>&1; echo "32 streams: $(tail -1 gpurun_out/step32.log | sed 's/.*mel ms/mel ms/')")
(timeout 300 python tools/profile_step.py --streams 8 --tokens 24 > gpurun_out/step8.log 2>&1; echo "8 streams: $(tail -1 gpurun_out/step8.log | sed 's/.*mel ms/mel ms/')")
(timeout 300 python tools/kbench.py > gpurun_out/kbench.txt 2>&1; sed -n 1,12p gpurun_out/kbench.txt)
(timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches32_r1.csv python tools/profile_step.py --streams 32 --tokens 12 --no-graph > gpurun_out/profile_step32.log 2>&1; echo "ncu list exit $?"; wc -l gpurun_out/launches32_r1.csv)
(timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"cross_attn_kernel" -s 40 -c 2 -o gpurun_out/prof_cross_r1 -f python tools/profile_step.py --streams 32 --tokens 3 --no-graph > gpurun_out/prof_cross.log 2>&1; echo "ncu cross exit $?")
(timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"gemm_tn_kernel" -s 1500 -c 6 -o gpurun_out/prof_gemm_dec_r1 -f python tools/profile_step.py --streams 32 --tokens 3 --no-graph > gpurun_out/prof_gemm_dec.log 2>&1; echo "ncu gemm dec exit $?")
(timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"gemm_tn_kernel" -s 12 -c 6 -o gpurun_out/prof_gemm_enc_r1 -f python tools/profile_step.py --streams 8 --tokens 3 --no-graph > gpurun_out/prof_gemm_enc.log 2>&1; echo "ncu gemm enc exit $?")
ls -la gpurun_out/*.ncu-rep

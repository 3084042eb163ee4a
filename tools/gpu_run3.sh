#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "max err|passed|failed|diverge|identical" gpurun_out/pytest_gpu.log | tail -40)
(WLB200_FUSED_ATTN=0 WLB200_SPLITK=0 timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "encoder or logits or generate" > gpurun_out/pytest_gpu_v1.log 2>&1; echo "pytest v1 exit $?"; tail -3 gpurun_out/pytest_gpu_v1.log)
(timeout 600 python bench.py --model small.en --streams 4 --steps 2 --warmup 1 --beam 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench small exit $?"; cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err)
(timeout 1500 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; echo "bench large exit $?"; cat gpurun_out/bench_large.json; tail -5 gpurun_out/bench_large.err)
(timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py --streams 8 --tokens 6 --no-graph > gpurun_out/profile_step.log 2>&1; echo "ncu list exit $?"; tail -2 gpurun_out/profile_step.log; wc -l gpurun_out/launches_r1.csv)
(timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"cross_attn_kernel|flash_attn_kernel" -c 4 -o gpurun_out/prof_attn_r1 python tools/profile_step.py --streams 8 --tokens 3 --no-graph > gpurun_out/prof_attn.log 2>&1; echo "ncu attn exit $?")
(timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"gemm_tn_kernel" -s 4 -c 6 -o gpurun_out/prof_gemm_r1 python tools/profile_step.py --streams 8 --tokens 3 --no-graph > gpurun_out/prof_gemm.log 2>&1; echo "ncu gemm exit $?")
ls -la gpurun_out/*.ncu-rep

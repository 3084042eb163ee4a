#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
K='test_generate_matches_oracle or test_generate_sampling_matches_oracle or test_large_v3_detect_language_and_align'
run() { name=$1; shift; (env "$@" timeout 400 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "$K" > gpurun_out/bisect_$name.log 2>&1; echo "== $name exit $?"; grep -E "^FAILED|passed|failed" gpurun_out/bisect_$name.log | tail -5); }
run default A=1
run pairflash WLB200_FA_SPLIT=0
run cgemm WLB200_CGEMM=1
run default2 A=1

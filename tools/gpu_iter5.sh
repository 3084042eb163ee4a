#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -s -k "encoder or full_size or small_en or jfk" > gpurun_out/pytest_enc.log 2>&1; echo "pytest encoder exit $?"; grep -E "passed|failed|Error|assert|max err|rel rms" gpurun_out/pytest_enc.log | tail -12)
(timeout 300 python tools/fa_ab.py --streams 16 --reps 3 2>&1 | tee gpurun_out/fa_ab2.txt | tail -4)
(timeout 400 ncu --profile-from-start off --clock-control none --set full --import-source on -k regex:flash_attn_kernel -s 3 -c 1 -o gpurun_out/prof_flash_r2c -f python tools/fa_ab.py --streams 16 --reps 1 > gpurun_out/prof_flash_r2c.log 2>&1; echo "ncu full exit $?"; python tools/summarize_ncu.py report gpurun_out/prof_flash_r2c.ncu-rep gpurun_out/prof_flash_r2c.md 2>&1 | grep -E "time_duration|tensor_cycles|registers" )

#!/bin/bash
# Round-2 evidence run (1 x B200): ncu launch list + one --set full capture per hot kernel.  Outputs land in gpurun_out/;
# tools/summarize_ncu.py turns them into the committed profiles/*.md.  Nothing printed under ncu is a bench number.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
P="python tools/profile_step.py"
(timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches32_r2.csv $P --streams 32 --tokens 12 --no-graph > gpurun_out/ncu_list32.log 2>&1; echo "list32 exit $?"; wc -l gpurun_out/launches32_r2.csv)
(timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches4_r2.csv $P --streams 4 --tokens 12 --no-graph > gpurun_out/ncu_list4.log 2>&1; echo "list4 exit $?"; wc -l gpurun_out/launches4_r2.csv)
cap() {  # name regex skip count streams tokens
  (timeout 600 $NCU --set full --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/prof_$1_r2 -f $P --streams $5 --tokens $6 --no-graph > gpurun_out/prof_$1.log 2>&1; echo "ncu $1 exit $?")
}
cap cross   "cross_attn_kernel"        40 2 32 3
cap decgemm "dec_gemm_kernel"          200 6 32 3
cap wgemm   "wgemm_kernel"             200 6 4 3
cap ln      "layernorm_update_kernel"  100 2 32 3
cap self    "self_attn_kernel"         40 2 32 6
cap search  "search_rows_kernel"       2 1 32 4
cap mel     "mel_stft_kernel"          0 1 32 2
cap flash   "flash_attn_kernel"        4 1 8 2
cap encgemm "gemm_tn_kernel"           12 6 8 2
ls -la gpurun_out/*_r2.ncu-rep

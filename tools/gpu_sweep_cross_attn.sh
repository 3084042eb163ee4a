#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "3 0" "4 0" "3 3" "4 3" "4 4" "2 2"; do
set -- $cfg
(WLB200_XA_STAGES=$1 WLB200_XA_NSPLIT=$2 timeout 200 python tools/profile_step.py --streams 32 --tokens 24 > gpurun_out/sweep_$1_$2.log 2>&1; echo "stages=$1 nsplit=$2: $(tail -1 gpurun_out/sweep_$1_$2.log | sed 's/.*enc ms/enc ms/')")
done

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 300 python tools/profile_step.py --streams 32 --tokens 24 > gpurun_out/step32.log 2>&1; echo "32 streams: $(tail -1 gpurun_out/step32.log | sed 's/.*mel ms/mel ms/')")
(WLB200_XA_NSPLIT=3 timeout 300 python tools/profile_step.py --streams 32 --tokens 24 > gpurun_out/step32_ns3.log 2>&1; echo "32 streams ns=3: $(tail -1 gpurun_out/step32_ns3.log | sed 's/.*mel ms/mel ms/')")
(WLB200_XA_NSPLIT=1 timeout 300 python tools/profile_step.py --streams 32 --tokens 24 > gpurun_out/step32_ns1.log 2>&1; echo "32 streams ns=1: $(tail -1 gpurun_out/step32_ns1.log | sed 's/.*mel ms/mel ms/')")
(WLB200_XA_STAGES=4 timeout 300 python tools/profile_step.py --streams 32 --tokens 24 > gpurun_out/step32_st4.log 2>&1; echo "32 streams stages=4: $(tail -1 gpurun_out/step32_st4.log | sed 's/.*mel ms/mel ms/')")
(timeout 300 python tools/profile_step.py --streams 8 --tokens 24 > gpurun_out/step8.log 2>&1; echo "8 streams: $(tail -1 gpurun_out/step8.log | sed 's/.*mel ms/mel ms/')")
(timeout 1500 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; echo "bench large exit $?"; cat gpurun_out/bench_large.json | cut -c1-400; tail -5 gpurun_out/bench_large.err)
(timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches32_r1.csv python tools/profile_step.py --streams 32 --tokens 12 --no-graph > gpurun_out/profile_step32.log 2>&1; echo "ncu list exit $?"; wc -l gpurun_out/launches32_r1.csv)

#!/bin/bash
# Round-2 closing run (1 x B200): smoke, the whole GPU suite, the bench lines committed under profiles/, ncu evidence for
# the kernels touched late in the round.  Nothing printed under ncu is a bench number.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log)
(timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|Error|assert|decode session|step rounds" gpurun_out/pytest_gpu.log | tail -10)
(timeout 600 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench (default flags) exit $?"; cut -c1-700 gpurun_out/bench_r2_n1.json; tail -3 gpurun_out/bench_r2_n1.err)
for s in 4 8 16; do
(timeout 300 python bench.py --streams $s --no-cpu-baseline > gpurun_out/bench_r2_n1_streams$s.json 2> gpurun_out/bench_s$s.err; echo "bench streams $s exit $?"; cut -c1-330 gpurun_out/bench_r2_n1_streams$s.json; tail -2 gpurun_out/bench_s$s.err)
done
(timeout 400 python bench.py --word-timestamps --no-cpu-baseline > gpurun_out/bench_r2_cfg4_word_timestamps.json 2> gpurun_out/bench_wt.err; echo "bench word-timestamps exit $?"; cut -c1-330 gpurun_out/bench_r2_cfg4_word_timestamps.json; tail -2 gpurun_out/bench_wt.err)
(timeout 200 python tools/enc_time.py --streams 32 --reps 3 2>&1 | tee gpurun_out/enc_time.txt | tail -3)
NCU="ncu --profile-from-start off --clock-control none"
(timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches32_r2b.csv python tools/profile_step.py --streams 32 --tokens 12 --no-graph > gpurun_out/ncu_list32b.log 2>&1; echo "list32 exit $?"; wc -l gpurun_out/launches32_r2b.csv)
(timeout 400 $NCU --set full --import-source on -k regex:flash_attn_kernel -s 3 -c 1 -o gpurun_out/prof_flash_r2d -f python tools/fa_ab.py --streams 16 --reps 1 > gpurun_out/prof_flash_r2d.log 2>&1; echo "ncu flash exit $?")

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?
echo "pytest (fused post-ops) exit $rc"; grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu.log | tail -12
if [ $rc -ne 0 ]; then
  WLB200_FUSE_POST=0 timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu_nofuse.log 2>&1; echo "pytest (unfused) exit $?"; tail -3 gpurun_out/pytest_gpu_nofuse.log
fi
for f in 1 0; do
(WLB200_FUSE_POST=$f timeout 300 python tools/profile_step.py --streams 32 --tokens 24 > gpurun_out/step32_f$f.log 2>&1; echo "32 streams fuse=$f: $(tail -1 gpurun_out/step32_f$f.log | sed 's/.*mel ms/mel ms/')")
(WLB200_FUSE_POST=$f timeout 300 python tools/profile_step.py --streams 4 --tokens 24 > gpurun_out/step4_f$f.log 2>&1; echo "4 streams fuse=$f: $(tail -1 gpurun_out/step4_f$f.log | sed 's/.*mel ms/mel ms/')")
done
(WLB200_TRACE=1 timeout 1500 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; echo "bench large exit $?"; cat gpurun_out/bench_large.json | cut -c1-900; tail -5 gpurun_out/bench_large.err)

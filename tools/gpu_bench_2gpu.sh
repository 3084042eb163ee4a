#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu exit $?"; tail -1 gpurun_out/bench_2gpu.json | cut -c1-1200; tail -3 gpurun_out/bench_2gpu.err)
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; echo "ref 2gpu exit $?"; tail -1 gpurun_out/bench_ref_2gpu.json | cut -c1-600; tail -3 gpurun_out/bench_ref_2gpu.err)

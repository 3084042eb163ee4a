#!/bin/bash
# Staged GPU validation: every stage runs in its own process under a timeout so that one hung or
# crashed kernel cannot take the rest of the run with it.  Logs land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
T=${STAGE_TIMEOUT:-420}
: > gpurun_out/summary.txt
run() {
  name=$1; shift
  timeout $T "$@" > gpurun_out/$name.log 2>&1
  rc=$?
  echo "== $name exit $rc" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed|error|max err|Error|diverge|identical|lengths" gpurun_out/$name.log | tail -${TAILN:-25} | tee -a gpurun_out/summary.txt
}
PT="python -m pytest tests/test_gpu_parity.py -q -s -p no:cacheprovider"
for stage in "$@"; do
  case $stage in
    gemm)      run gemm $PT -k gemm ;;
    mel)       run mel $PT -k mel ;;
    enc_simt)  WLB200_GEMM_SIMT=1 run enc_simt $PT -k "encoder or logits" ;;
    enc)       run enc $PT -k "encoder or logits" ;;
    gen_simt)  WLB200_GEMM_SIMT=1 run gen_simt $PT -k "generate or detect or align or slot or transcribe" ;;
    gen)       run gen $PT -k "generate or detect or align or slot or transcribe" ;;
    all)       run all python -m pytest tests -q -m gpu -p no:cacheprovider ;;
    *)         echo "unknown stage $stage" ;;
  esac
done

#!/usr/bin/env bash
# compute-sanitizer passes over the kernel self-tests (run on the GPU box; slow -- one check per tool).
#   tools/sanitize.sh memcheck fprop_3x3        tools/sanitize.sh racecheck bn_relu_res
# memcheck: out-of-bounds / misaligned global+shared accesses; racecheck: shared-memory hazards between the
# producer / MMA / epilogue warps; synccheck: illegal barrier use.  Results go to gpurun_out/sanitize_<tool>_<check>.log
set -u
tool=${1:-memcheck}; check=${2:-fprop_3x3}
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool "$tool" --print-limit 20 python tools/gpu_selftest.py --run "$check" \
  > "gpurun_out/sanitize_${tool}_${check}.log" 2>&1
tail -5 "gpurun_out/sanitize_${tool}_${check}.log"

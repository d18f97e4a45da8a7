#!/usr/bin/env bash
# One `ncu --set full` capture of a single kernel launch, named by a regex over the demangled kernel name.
#   tools/ncu_kernel.sh <out-name> '<regex>' <skip> -- <python command ...>
# e.g. 3x3 fprop of ResNet-50 shape 15 (template args print as "(int)256, (int)0"):
#   tools/ncu_kernel.sh prof_fprop3x3 'conv_gemm_kernel<\(int\)[0-9]+, \(int\)0>' 3 -- python tools/bench_conv.py --shapes 15 --iters 1
# Writes gpurun_out/<out-name>.ncu-rep; summarise with `ncu -i ... --page raw --csv` / `--page source --csv` into profiles/.
set -u
out=$1; regex=$2; skip=$3; shift 3
[ "${1:-}" = "--" ] && shift
mkdir -p gpurun_out
exec ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:${regex}" --launch-skip "${skip}" -c 1 -o "gpurun_out/${out}" -f "$@"

#!/usr/bin/env bash
# One `ncu --set full` capture of the 3x3 wgrad tcgen05 kernel (ResNet-50 shape 256->256 @14x14, batch 256).
# Output: gpurun_out/prof_conv_wgrad3x3.ncu-rep (summarise into profiles/ with `ncu -i ... --page raw --csv`).
mkdir -p gpurun_out
exec ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:conv_gemm_kernel<\(int\)[0-9]+, \(int\)1>' -c 1 -o gpurun_out/prof_conv_wgrad3x3 -f \
  python tools/bench_conv.py --shapes "${1:-15}" --iters 1 --out gpurun_out/conv_one.json

#!/usr/bin/env bash
# One-call GPU validation of the working tree (run through gpurun): gate -> pytest -m gpu -> smoke -> bench (JSON to
# gpurun_out/bench1.json) [-> per-kernel launch list + roofline tables with --profile].
#   gpurun --timeout 600 -- 'tools/gpu_validate.sh'            (~2.5 min of box time)
#   gpurun --timeout 900 -- 'tools/gpu_validate.sh --profile'  (+ conv/BN shape tables and the launch list, ~6 min)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools/gpu_gate.sh || exit 99
rc=0
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 || rc=1
tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2 || rc=1
timeout 180 python bench.py --steps 20 --warmup 5 > gpurun_out/bench1.json 2> gpurun_out/bench1.err || rc=1
cut -c1-320 gpurun_out/bench1.json
if [ "${1:-}" = "--profile" ]; then
  timeout 200 python tools/bench_conv.py > gpurun_out/bench_conv.log 2>&1; tail -1 gpurun_out/bench_conv.log
  timeout 120 python tools/bench_bn.py > gpurun_out/bench_bn.log 2>&1; tail -1 gpurun_out/bench_bn.log
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-e2e > gpurun_out/ncu_bench.log 2>&1
  python tools/roofline.py --conv gpurun_out/conv_shapes.json --launches gpurun_out/launches.csv --out gpurun_out/ROOFLINE.md > /dev/null 2>&1
fi
exit $rc

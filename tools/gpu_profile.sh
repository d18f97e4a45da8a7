#!/usr/bin/env bash
# One-GPU profiling session (run through gpurun, ~9 min of box time):
#   gate -> pytest -m gpu -> bench (graph on/off) -> conv/BN shape tables vs cuDNN -> launch list -> roofline
#   -> one `ncu --set full` capture per kernel family (halo conv, depthwise, attention, loss, optimizer).
# Everything lands in gpurun_out/; copy what should be judged into profiles/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools/gpu_gate.sh || exit 99
rc=0
timeout 150 python tools/bench_conv.py --ours-only --shapes 1 --out gpurun_out/conv_halo.json 2>&1 | tail -3
for g in on off; do
  timeout 180 python bench.py --steps 20 --warmup 5 --graph $g > gpurun_out/bench1_graph_$g.json 2> gpurun_out/bench1_graph_$g.err || rc=1
  cut -c1-260 gpurun_out/bench1_graph_$g.json
done
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 || rc=1
tail -3 gpurun_out/pytest_gpu.log
timeout 240 python tools/bench_conv.py > gpurun_out/bench_conv.log 2>&1; tail -1 gpurun_out/bench_conv.log
timeout 120 python tools/bench_bn.py > gpurun_out/bench_bn.log 2>&1; tail -1 gpurun_out/bench_bn.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-e2e --graph off > gpurun_out/ncu_bench.log 2>&1
python tools/roofline.py --conv gpurun_out/conv_shapes.json --launches gpurun_out/launches.csv --out gpurun_out/ROOFLINE.md > /dev/null 2>&1
# name | kernel regex | launch-skip | arch
while IFS='|' read -r name regex skip arch; do
  [ -z "$name" ] && continue
  timeout 200 tools/ncu_kernel.sh "$name" "$regex" "$skip" -- python bench.py --arch "$arch" --steps 1 --warmup 3 --skip-e2e --graph off \
    > "gpurun_out/$name.log" 2>&1
  if [ -f "gpurun_out/$name.ncu-rep" ]; then
    python tools/ncu_summary.py "gpurun_out/$name.ncu-rep" 20 > "gpurun_out/ncu_$name.txt" 2>/dev/null
    head -8 "gpurun_out/ncu_$name.txt" | cut -c1-150
  else
    echo "no capture for $name"; tail -2 "gpurun_out/$name.log"
  fi
done <<'LIST'
prof_halo_fprop|conv3x3_halo_kernel|8|resnet50
prof_stem_tail_fwd|bn_relu_pool_fwd_strip_kernel|2|resnet50
prof_stem_tail_bwd|bn_relu_pool_bwd_strip_kernel|3|resnet50
prof_ce_topk|ce_topk_kernel|2|resnet50
prof_sgd_local|sgd_local_kernel|2|resnet50
prof_dw_fprop|dw_fprop_fast_kernel|20|efficientnet_b0
prof_dw_wgrad|dw_wgrad_fast_kernel|20|efficientnet_b0
prof_attn_fwd|attn_fwd_kernel|6|botnet50
prof_attn_bwd_dq|attn_bwd_dq_kernel|6|botnet50
prof_se_bwd_w|se_gate_bwd_weights_kernel|20|regnety_160
LIST
# keep the merge under the 64 MiB cap: reports other than the halo one are summarised above and dropped
for f in gpurun_out/prof_*.ncu-rep; do case "$f" in *halo*|*attn_fwd*) ;; *) rm -f "$f";; esac; done
du -sh gpurun_out | cut -f1
exit $rc

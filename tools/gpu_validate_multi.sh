#!/usr/bin/env bash
# Multi-GPU validation (run through `gpurun --gpus N`):  tools/gpu_validate_multi.sh N [--archs]
#   gate -> tools/multigpu_check.py at N ranks -> bench.py with the captured step on and off (BENCH_MODES="on" for one)
#   [-> the other families].
set -u
cd "$(dirname "$0")/.."
n=${1:-2}
mkdir -p gpurun_out
tools/gpu_gate.sh || exit 99
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29601 tools/multigpu_check.py > gpurun_out/multigpu$n.log 2>&1; echo "multigpu_check rc=$?"
grep -E "^PASS|^FAIL" gpurun_out/multigpu$n.log | cut -c1-300
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]),
          round(d.get("e2e_uint8_input", {}).get("value", 0)), d["clocks"]["reasons"], json.dumps(d.get("multi_gpu_attribution")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
for g in ${BENCH_MODES:-on off}; do
  f=gpurun_out/bench_ours_${n}gpu_graph_$g
  timeout 240 $TR --master-port 29602 bench.py --gpus $n --steps 20 --warmup 5 --graph $g > $f.json 2> $f.err
  show $f.json; tail -1 $f.err | cut -c1-200
done
if [ "${2:-}" = "--archs" ]; then
  for a in botnet50 regnety_160 efficientnet_b0; do
    f=gpurun_out/bench_${a}_${n}gpu
    timeout 200 $TR --master-port 29603 bench.py --arch $a --gpus $n --steps 20 --warmup 5 > $f.json 2> $f.err
    show $f.json; tail -1 $f.err | cut -c1-200
  done
fi

#!/usr/bin/env bash
# First step of every GPU session: fail fast (exit 99) if the box cannot even initialise CUDA within 90 s,
# so a wedged device costs one minute instead of the whole session.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,temperature.gpu,memory.used --format=csv > gpurun_out/gate_smi.txt 2>&1
if ! timeout 90 python -c "import torch; x=torch.ones(1024,device='cuda'); torch.cuda.synchronize(); print('cuda ok', torch.cuda.get_device_name(0), float(x.sum()))"; then
  echo "GATE FAILED: CUDA init/compute did not finish in 90s"; cat gpurun_out/gate_smi.txt; exit 99
fi

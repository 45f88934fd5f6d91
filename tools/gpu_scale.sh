#!/bin/bash
# bench at N GPUs under torchrun, as the driver launches it
N=${1:-2}; R=${2:-r01}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_${R}_n$N.json 2> gpurun_out/scale_${R}_n$N.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_${R}_n$N.json 2> gpurun_out/scale_${R}_n$N.err
fi
echo "rc=$?"; cat gpurun_out/scale_${R}_n$N.json | cut -c1-1200; tail -5 gpurun_out/scale_${R}_n$N.err

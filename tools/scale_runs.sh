#!/bin/bash
# Multi-GPU measurement set of round 2 (run on an 8-GPU box: gpurun --gpus 8 -- bash tools/scale_runs.sh).
# BASELINE cfg 4 as written (8192 images per step, strong scaling), cfg 5 (training step at 1/2/4/8 GPUs), the in-process sharded
# driver and the host-side copy ceiling.  Every JSON line lands in gpurun_out/ (copied to profiles/ by hand).
set -u
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29700
for N in 1 2 4 8; do
  port=$((port+1))
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --total-batch 8192 --steps 8 --warmup 3 --preheat 1 --no-cpu-baseline --no-fp32-grade --no-api --check-images 32 > $OUT/r02_strong8192_${N}gpu.json 2> $OUT/r02_strong8192_${N}gpu.log
  else
    $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --total-batch 8192 --steps 8 --warmup 3 --preheat 1 --no-cpu-baseline --no-fp32-grade --no-api --check-images 32 > $OUT/r02_strong8192_${N}gpu.json 2> $OUT/r02_strong8192_${N}gpu.log
  fi
  echo "strong8192 N=$N rc=$?"
done
for N in 1 2 4 8; do
  port=$((port+1))
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --workload train --steps 20 --warmup 3 > $OUT/r02_train_${N}gpu.json 2> $OUT/r02_train_${N}gpu.log
  else
    $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --workload train --steps 20 --warmup 3 > $OUT/r02_train_${N}gpu.json 2> $OUT/r02_train_${N}gpu.log
  fi
  echo "train N=$N rc=$?"
done
python bench.py --workload sharded --gpus 8 --steps 8 > $OUT/r02_sharded_8gpu.json 2> $OUT/r02_sharded_8gpu.log; echo "sharded rc=$?"
python bench.py --workload sharded --gpus 8 --total-batch 8192 --steps 5 > $OUT/r02_sharded_8gpu_8192.json 2> $OUT/r02_sharded_8gpu_8192.log; echo "sharded 8192 rc=$?"
python tools/host_ceiling_probe.py > $OUT/r02_host_ceiling.log 2>&1; echo "ceiling rc=$?"
python tools/host_ceiling_probe.py --numa > $OUT/r02_host_ceiling_numa.log 2>&1
port=$((port+1))
$TR --nproc-per-node 8 --master-port $port bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-grade --check-images 32 > $OUT/r02_weak_8gpu.json 2> $OUT/r02_weak_8gpu.log; echo "weak8 rc=$?"
nvidia-smi topo -m > $OUT/r02_topo.txt 2>&1
tail -2 $OUT/r02_host_ceiling.log

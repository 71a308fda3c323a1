#!/bin/bash
# Second multi-GPU set of round 2 (after the second-generation fused kernel and the tensor-core training kernels):
#   gpurun --gpus 8 -- bash tools/scale_runs_b.sh [N ...]     -> gpurun_out/r02b_*.json (copied to profiles/ by hand)
set -u
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29800
for N in ${@:-2 4 8}; do
  port=$((port+1))
  $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --workload train --steps 20 --warmup 3 > $OUT/r02b_train_${N}gpu.json 2> $OUT/r02b_train_${N}gpu.log
  echo "train N=$N rc=$?"
done
if [ "${STRONG:-0}" = "1" ]; then
port=$((port+1))
$TR --nproc-per-node 8 --master-port $port bench.py --gpus 8 --total-batch 8192 --steps 8 --warmup 3 --preheat 1 --no-cpu-baseline --no-fp32-grade --no-api --check-images 32 > $OUT/r02b_strong8192_8gpu.json 2> $OUT/r02b_strong8192_8gpu.log
echo "strong8192 N=8 rc=$?"
fi
python - "$@" <<'P'
import json, sys
for n in (sys.argv[1:] or ['2','4','8']):
    f=f'r02b_train_{n}gpu'
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d.get('comm') and {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['comm'].items() if k!='note'})
    except Exception as e:
        print(f, 'ERR', e)
P

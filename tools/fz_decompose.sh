#!/bin/bash
# decomposition of conv_fused_kernel by its debug bits (timing only; results are garbage)
#   1 no activation store, 2 no fc1-tile store, 4 empty epilogue, 8 no second-block MMAs, 16 first-block teams idle,
#   128 second-block MMAs predicated off
for dbg in ${FZ_BITS:-0 16 20 4 8 24 17 19 12}; do
  BCAD_DEBUG_SKIP_STORES=$dbg python bench.py --steps 50 --warmup 5 --preheat 1 --no-check --no-cpu-baseline --no-fp32-grade --only-value --refine-margin 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=[x for x in d['kernels'] if 'conv01' in x['kernel']][0]
print('debug=$dbg', 'step ms %.4f'%d['ms_per_step'], 'fused ms', k['ms'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
done

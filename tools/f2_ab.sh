#!/bin/bash
# same-box A/B of the fused conv kernels: sustained (2 s pre-heat) and burst (idle 3 s, no pre-heat) timings per build / variant
# (build the variants first: tools/build_variant.sh t4e4 "-DF2_TEAM_WARPS=4 -DF2_EPI_WARPS=4"; tools/build_variant.sh t4e8 "-DF2_TEAM_WARPS=4 -DF2_EPI_WARPS=8";
#  the default build is t4e4 since r02 -- run "t8e4" through a variant built with -DF2_TEAM_WARPS=8)
D=$PWD/vision-xai-breast-cancer-cad_b200
run() {  # label, env assignments...
  label=$1; shift
  for mode in sustained burst; do
    if [ $mode = burst ]; then sleep 3; pre=0; steps=20; else pre=2; steps=50; fi
    env "$@" python bench.py --steps $steps --warmup 5 --preheat $pre --no-check --no-cpu-baseline --no-fp32-grade --only-value --refine-margin 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=[x for x in d['kernels'] if 'conv01' in x['kernel']][0]
print('$label $mode', 'step ms %.4f'%d['ms_per_step'], 'fused ms', k['ms'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
  done
}
for rep in 1 2; do
run v1 BCAD_FUSED_V1=1
run default BCAD_X=1
run t8e4 BCAD_LIB=$D/libbcad_t8e4.so
run t4e4 BCAD_LIB=$D/libbcad_t4e4.so
run t4e8 BCAD_LIB=$D/libbcad_t4e8.so
done

cd /root/repo
D=$PWD/vision-xai-breast-cancer-cad_b200
for at in 3 300 1500; do
echo "== t4e4 trace lib, bench mode, trace of launch $at"
BCAD_F2_TRACE_AT=$at BCAD_LIB=$D/libbcad_t4e4tr.so timeout 200 python bench.py --steps 50 --warmup 5 --preheat 2 --no-check --no-cpu-baseline --no-fp32-grade --only-value --refine-margin 0 2>&1 >/dev/null | grep "f2_trace issuer\|f2_trace team\|per-CTA issuer"
done > gpurun_out/r02q_clock.log 2>&1
cat gpurun_out/r02q_clock.log | cut -c1-400

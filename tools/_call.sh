cd /root/repo
CMD="python bench.py --steps 3 --warmup 3 --preheat 0 --no-check --no-cpu-baseline --no-fp32-grade --only-value"
$CMD > gpurun_out/r02t_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02t_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv_fused2|tail_fused|fc_splitk|dense_head" -s 16 -c 4 -o gpurun_out/r02_all -f $CMD > gpurun_out/r02t_ncu2.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/r02t_ncu2.log
ls -la gpurun_out/r02_all.ncu-rep gpurun_out/r02_launches.csv

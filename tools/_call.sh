cd /root/repo
timeout 600 python -m pytest tests/test_gpu_training.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --workload train > gpurun_out/r03d_train.json 2> gpurun_out/r03d_train.err; echo "rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r03d_train.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'loss', d['e2e'].get('last_mean_loss'))
for k in d['kernels']: print('   ', k['kernel'], k['ms'])
P

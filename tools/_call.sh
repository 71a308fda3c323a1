cd /root/repo
timeout 900 python tools/tc_train_probe.py 2>&1 | tee gpurun_out/r03a_probe.log
timeout 600 python -m pytest tests/test_gpu_training.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --workload train --train-kernels tensor > gpurun_out/r03a_train_tensor.json 2> gpurun_out/r03a_train_tensor.err; echo "rc=$?"
python - <<'P'
import json
for f in ['r03a_train_tensor']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'loss', d['e2e'].get('last_mean_loss'))
        for k in d['kernels']: print('   ', k['kernel'], k['ms'])
    except Exception as e:
        print(f, 'ERR', e)
P

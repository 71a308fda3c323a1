cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r03e_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r03e_tests.log
timeout 300 python bench.py --workload train > gpurun_out/r03e_train.json 2> gpurun_out/r03e_train.err; echo "train rc=$?"
timeout 300 python bench.py --workload train --train-kernels fp32 > gpurun_out/r03e_train_fp32.json 2> gpurun_out/r03e_train_fp32.err; echo "train fp32 rc=$?"
timeout 600 python bench.py > gpurun_out/r03e_bench.json 2> gpurun_out/r03e_bench.err; echo "bench rc=$?"
python - <<'P'
import json
for f in ['r03e_train','r03e_train_fp32','r03e_bench']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value']), 'ms', round(d.get('ms_per_step',0),4), 'e2e', d.get('e2e',{}).get('value'), d.get('clocks'))
        print('  kernels', [(k['kernel'],k['ms']) for k in d['kernels']])
        if f=='r03e_bench':
            print('  roofline', {k:d['roofline'][k] for k in ('kernel','ms','achieved','frac','frac_burst','traffic')})
            print('  check', {k:d['check'][k] for k in ('images','class_mismatches','max_logit_err','max_heat_err','maps_out_of_tolerance')})
    except Exception as e:
        print(f, 'ERR', e)
P

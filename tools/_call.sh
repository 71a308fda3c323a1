cd /root/repo
timeout 900 python tools/tc_train_probe.py 2>&1 | tee gpurun_out/r02w_probe.log

cd /root/repo
python tools/dev_determinism.py > gpurun_out/r2_det.log 2>&1

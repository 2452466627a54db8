cd /root/repo
python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -12 > gpurun_out/r2_t7.log

cd /root/repo
timeout 60 python -m pytest tests/test_streaming.py tests/test_gpu_parity.py -q -m gpu -x -k "stream or apply_model_matches or separator_front_door or errors_are_loud" 2>&1 | tail -5 > gpurun_out/r2_last.log

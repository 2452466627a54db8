cd /root/repo
python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py -q -m gpu -x -k "not dconv and not conv0" 2>&1 | tail -15 > gpurun_out/r2_t1.log
python tools/dev_modes.py strict,bf16,tf32x3 > gpurun_out/r2_modes1.log 2>&1
for m in strict bf16 tf32; do python -m demucs_b200.perf --batch 16 --mode $m --top 45 > gpurun_out/r2_perf_$m.txt 2>&1; done

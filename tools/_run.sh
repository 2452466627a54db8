cd /root/repo
python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_t3.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_strict_b16.json 2> gpurun_out/r2_bench_strict_b16.err
python bench.py --steps 3 --warmup 3 --batch 64 --no-cpu > gpurun_out/r2_bench_strict_b64.json 2> gpurun_out/r2_bench_strict_b64.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2>&1

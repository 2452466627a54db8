cd /root/repo
timeout 300 python -m pytest tests/test_hdemucs.py -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_h1.log
timeout 300 python bench.py --config hdemucs_mmi --steps 3 --warmup 3 > gpurun_out/r2_bench_hdemucs.json 2> gpurun_out/r2_bench_hdemucs.err

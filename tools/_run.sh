cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29621 tools/dist_check.py strict > gpurun_out/r2_dist8.log 2>&1
timeout 400 $TR --master-port 29622 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
timeout 400 $TR --master-port 29623 bench.py --gpus 8 --steps 5 --warmup 3 --batch 4 --no-weak > gpurun_out/r2_bench_n8_b4.json 2> gpurun_out/r2_bench_n8_b4.err
timeout 500 $TR --master-port 29624 bench.py --gpus 8 --steps 3 --warmup 3 --config 6s_10min > gpurun_out/r2_bench_6s_n8.json 2> gpurun_out/r2_bench_6s_n8.err
timeout 500 $TR --master-port 29625 bench.py --gpus 8 --steps 3 --warmup 3 --config ft_10min > gpurun_out/r2_bench_ft_n8.json 2> gpurun_out/r2_bench_ft_n8.err

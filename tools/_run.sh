cd /root/repo
CMD="python -m demucs_b200.perf --batch 16 --mode strict --top 10"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active
$CMD > gpurun_out/plain_perf.log 2>&1 && ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_forward_b16_strict.csv $CMD > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"persist_kernel<32, 256, 2>|attention_b16" -s 94 -c 8 -o gpurun_out/r02_full_gemm_attn $CMD > gpurun_out/ncu_b.log 2>&1
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu"
$BCMD > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 800 --csv --log-file gpurun_out/r02_ncu_launches_bench_step.csv $BCMD > gpurun_out/ncu_c.log 2>&1

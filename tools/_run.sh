cd /root/repo
CMD="python -m demucs_b200.perf --batch 16 --mode strict --top 5"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:.*(persist_kernel<\(int\)32, \(int\)256, \(int\)2>|attention_b16_kernel).*' -s 94 -c 7 -o gpurun_out/r02_full_gemm_attn $CMD > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/ | tail -3 >> gpurun_out/ncu_d.log

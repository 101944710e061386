timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fused_fwd' -s 8 -c 1 -o gpurun_out/prof_ff python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ff.log 2>&1
tail -3 gpurun_out/ncu_ff.log

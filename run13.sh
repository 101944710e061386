timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fused_bwd|fused_fwd' -s 8 -c 2 -o gpurun_out/prof_f2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1
tail -2 gpurun_out/ncu_f2.log

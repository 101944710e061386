mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'gemm_pair|folded_seed' -s 40 -c 14 -f -o gpurun_out/r2_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/r2_prof.ncu-rep
ncu -i gpurun_out/r2_prof.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size 2>/dev/null | cut -c1-700 | head -30

mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'critic_pre|critic_post|pose_fwd|pose_bwd|adam_params|loss_finish|folded_seed' -s 30 -c 14 -f -o gpurun_out/r2_prof_small python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log | cut -c1-200
ncu -i gpurun_out/r2_prof_small.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__inst_executed.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin))
hdr=r[0]
ki=hdr.index('Kernel Name')
cols=[i for i,h in enumerate(hdr) if '__' in h]
seen=set()
for row in r[2:]:
    n=row[ki].split('(')[0]
    if n in seen: continue
    seen.add(n)
    print(n)
    for i in cols: print('   ',hdr[i][:80], row[i], r[1][i])
"

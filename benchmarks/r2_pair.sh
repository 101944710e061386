# CTA-pair GEMM: correctness + timing per dispatch mode, then the step with and without it
for mode in 1 0 2; do
  JRR_GEMM_PAIR=$mode timeout 300 python benchmarks/gemm_pair_check.py > gpurun_out/r2_pair_mode$mode.jsonl 2> gpurun_out/r2_pair_mode$mode.err || echo "mode $mode FAILED rc=$?"
  tail -3 gpurun_out/r2_pair_mode$mode.err
  cat gpurun_out/r2_pair_mode$mode.jsonl
done

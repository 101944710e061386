timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -4 > gpurun_out/r2_tests_tail.txt; cat gpurun_out/r2_tests_tail.txt
for cfg in "JRR_CRITIC_POST_FUSED=0" "JRR_CRITIC_POST_FUSED=1"; do
  env $cfg timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/r2_tmp.json 2> gpurun_out/r2_tmp.err || { echo "$cfg FAILED"; tail -3 gpurun_out/r2_tmp.err; }
  python - "$cfg" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r2_tmp.json").read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], [(k['name'][:14],k['ms']) for k in d['kernels']], d['quality']['oracle_one_step_rel_err'], d['quality']['mpjpe_after_mm'])
PY
done

# folded-path GEMMs on CTA pairs: parity tests, then the step per tile-width / off
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -k "folded or bench_config or refit or gemm or multi_step or loss_history" 2>&1 | tail -4
for cfg in "JRR_FOLD_TS=0" "JRR_FOLD_BN=256" "JRR_FOLD_BN=192" "JRR_FOLD_BN=128"; do
  env $cfg timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/r2_fold_tmp.json 2> gpurun_out/r2_fold_tmp.err || { echo "$cfg FAILED"; tail -3 gpurun_out/r2_fold_tmp.err; }
  python - "$cfg" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r2_fold_tmp.json").read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], [(k['name'][:14],k['ms']) for k in d['kernels']], d['quality']['oracle_one_step_rel_err'])
PY
done

# quick A/B: GPU tests (optional: TESTS=1), then bench.py per environment setting given as arguments ("-" = defaults)
if [ "$TESTS" = "1" ]; then timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -6 > gpurun_out/r2_tests_tail.txt; cat gpurun_out/r2_tests_tail.txt; fi
for cfg in "$@"; do
  if [ "$cfg" = "-" ]; then envs=""; else envs="$cfg"; fi
  env $envs timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/r2_tmp.json 2> gpurun_out/r2_tmp.err || { echo "$cfg FAILED"; tail -3 gpurun_out/r2_tmp.err; }
  python - "$cfg" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r2_tmp.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value']), d['ms_per_step'], round(d['e2e']['value']), [(k['name'][:14],k['ms']) for k in d['kernels']], d['quality']['mpjpe_after_mm'], round(d['other_loss_path']['value']))
PY
done

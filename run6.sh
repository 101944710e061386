timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -6
for ff in 0 1; do JRR_FUSED_FWD=$ff timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_ff$ff.json 2> gpurun_out/bench_ff$ff.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_ff$ff.json'))
print("fused=$ff", d['value'], d['ms_per_step'], d['e2e']['value'], [(k['name'],k['ms']) for k in d['kernels'][:6]], d['quality'])
PY
tail -2 gpurun_out/bench_ff$ff.err; done

python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 > gpurun_out/t_gpu.log; tail -8 gpurun_out/t_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1b.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['kernels']: print(k)
print(d['quality'], d['refit_ms'], d['clocks'])
PY
tail -3 gpurun_out/bench_r1b.err

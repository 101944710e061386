# GPU tests, then the bench line
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -4 | tee gpurun_out/r2_gpu_tests_tail.txt
timeout 600 python bench.py > gpurun_out/r2_bench_pair.json 2> gpurun_out/r2_bench_pair.err; tail -2 gpurun_out/r2_bench_pair.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_pair.json").read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline'])
print([(k['name'][:14],k['ms'],k.get('frac')) for k in d['kernels']])
print(d['whole_step'], d['other_loss_path']['value'], d['clocks'])
PY

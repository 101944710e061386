mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -x -k "gemm or critic or single_step or deterministic" 2>&1 | tail -3
timeout 200 python benchmarks/gemm_prof.py 2>&1 | tail -4
JRR_GEMM_PROBE=32 timeout 200 python benchmarks/gemm_prof.py 2>&1 | tail -4 | head -1
timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline 2> gpurun_out/r2_bench_c.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['refit_ms'], d['roofline']['frac'], d['whole_step']['tensor_frac_3xtf32'])
print([(k['name'][:18],k['ms'],k.get('frac')) for k in d['kernels']])"
tail -3 gpurun_out/r2_bench_c.err

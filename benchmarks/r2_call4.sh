mkdir -p gpurun_out
timeout 200 python benchmarks/gemm_prof.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -s -x -k "folded_operator or follows_the_refit or refit_matches or gemm" 2>&1 | grep -E "^\[|passed|failed|assert|Error|error" | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline 2> gpurun_out/r2_bench_c.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['refit_ms'])"
tail -3 gpurun_out/r2_bench_c.err

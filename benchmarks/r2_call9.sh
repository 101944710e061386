mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline 2> gpurun_out/r2_bench_c.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['refit_ms'], d['other_loss_path']['value'])"
tail -3 gpurun_out/r2_bench_c.err

python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -2 gpurun_out/r2_bench.err
python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/r2_bench_k20.json 2>> gpurun_out/r2_bench.err
python - <<'PY'
import json
for f in ("r2_bench", "r2_bench_k20"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d['value']), d['ms_per_step'], "e2e", round(d['e2e']['value']), d['e2e']['ms_per_step'], d['whole_step']['tensor_frac_3xtf32'], d['refit_ms'], d['clocks'])
    if d.get('secondary'):
        print(json.dumps(d['secondary']['c5_smpl_module']))
        print(json.dumps(d['secondary']['silhouette_term']))
        print(d['secondary']['c3_strong']['seconds'])
PY

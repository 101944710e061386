timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -s 2>&1 | grep -E "^\[|MPJPE|refit|gradient rel|tcgen05 3xTF32|passed|failed|rel err|max \|" > gpurun_out/r2_gpu_tests.txt; tail -1 gpurun_out/r2_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -1 gpurun_out/r2_smoke.txt | cut -c1-200
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -2 gpurun_out/r2_bench.err
python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/r2_bench_k20.json 2>> gpurun_out/r2_bench.err
python bench.py --loss-path vertex --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_vertex.json 2>> gpurun_out/r2_bench.err
python - <<'PY'
import json
for f in ("r2_bench", "r2_bench_k20", "r2_bench_vertex"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d['value']), d['ms_per_step'], "e2e", round(d['e2e']['value']), d['e2e']['ms_per_step'], d['whole_step']['tensor_frac_3xtf32'], d['refit_ms'], d['clocks']['reasons'])
PY

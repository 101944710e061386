# Regenerates everything under profiles/ on a B200 (run through gpurun from the repo root):
#   gpurun --timeout 2400 -- "bash benchmarks/gpu_validate.sh r2"   then   python profiles/summarize.py r2
T=${1:-r2}
P=${2:-all}      # 1 = tests, bench lines, launch list, ncu captures of the step; 2 = module-path captures, GEMM checks, breakdown, memcheck
# (gpurun brings back at most 64 MiB per call: the three ncu reports together exceed that, so run the two parts as two calls)
mkdir -p gpurun_out
if [ "$P" != "2" ]; then
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s 2>&1 | grep -E "^\[|MPJPE|refit|gradient rel|tcgen05 3xTF32|passed|failed|rel err|max \|" > gpurun_out/${T}_gpu_tests.txt; tail -1 gpurun_out/${T}_gpu_tests.txt
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/${T}_bench_k20.json 2>> gpurun_out/${T}_bench.err      # the driver's command line
python bench.py --loss-path vertex --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_vertex.json 2>> gpurun_out/${T}_bench.err
python bench.py --regressor shipped --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench_shipped.json 2>> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
python - $T <<'PY'
import json,sys
T=sys.argv[1]
for f in ("bench","bench_vertex","bench_shipped"):
    d=json.loads(open(f"gpurun_out/{T}_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['gpu_launches'], d['roofline'].get('kernel'), d['roofline'].get('frac'), d['whole_step'].get('tensor_frac_3xtf32'))
    print([(k['name'][:14],k['ms'],k.get('frac')) for k in d['kernels']])
print(open(f"gpurun_out/{T}_bench_reference.json").read()[:300])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm_pair|folded_seed|critic_pre|critic_post|pose_fwd|pose_bwd|adam_params' -s 60 -c 20 -f -o gpurun_out/${T}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fused_bwd|fused_fwd' -s 30 -c 4 -f -o gpurun_out/${T}_prof_vertex python bench.py --loss-path vertex --steps 3 --warmup 3 --no-cpu-baseline --no-secondary >> gpurun_out/ncu_full.log 2>&1
fi
if [ "$P" != "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:'fused_fwd|fused_bwd|smpl_small' -c 13 -f -o gpurun_out/${T}_prof_module python benchmarks/module_calls.py >> gpurun_out/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_module_launches.csv python benchmarks/module_calls.py > gpurun_out/${T}_module.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sil_|small_bwd' -c 12 -f -o gpurun_out/${T}_prof_sil python benchmarks/sil_calls.py >> gpurun_out/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_silhouette_launches.csv python benchmarks/sil_calls.py > gpurun_out/${T}_sil.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; tail -1 gpurun_out/${T}_smoke.txt
ls -la gpurun_out/${T}_prof*.ncu-rep; du -sh gpurun_out
timeout 300 python benchmarks/gemm_pair_check.py > gpurun_out/${T}_gemm_pair_check.jsonl 2>/dev/null
timeout 300 python benchmarks/gemm_prof.py > gpurun_out/${T}_gemm_role_stamps.jsonl 2>/dev/null
for s in 0 1 2 3; do JRR_DEBUG_SKIP=$s timeout 200 python benchmarks/step_breakdown.py; done > gpurun_out/${T}_step_breakdown.jsonl 2>/dev/null
JRR_OVERLAP_CRITIC=0 timeout 200 python benchmarks/step_breakdown.py >> gpurun_out/${T}_step_breakdown.jsonl 2>/dev/null
compute-sanitizer --tool memcheck --print-limit 20 python benchmarks/sanitize.py 2>&1 | grep -E "COMPUTE-SANITIZER|ERROR SUMMARY|Invalid|sanitizer workload|at 0x|Error" | head -40 > gpurun_out/${T}_memcheck.txt; tail -2 gpurun_out/${T}_memcheck.txt
fi

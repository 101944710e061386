# Regenerates everything under profiles/ on a B200 (run through gpurun from the repo root):
#   gpurun --timeout 1800 -- "bash benchmarks/gpu_validate.sh"   then copy gpurun_out/r1_* to profiles/
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4 | tee gpurun_out/r1_gpu_tests_tail.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s 2>&1 | grep -E "^\[|MPJPE|refit|gradient rel|tcgen05 3xTF32|passed|failed" > gpurun_out/r1_gpu_tests.txt
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; tail -2 gpurun_out/r1_bench.err
python bench.py --loss-path vertex --no-cpu-baseline > gpurun_out/r1_bench_vertex.json 2>> gpurun_out/r1_bench.err
python bench.py --regressor shipped --no-cpu-baseline > gpurun_out/r1_bench_shipped.json 2>> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench.err
python - <<'PY'
import json
for f in ("r1_bench","r1_bench_vertex","r1_bench_shipped"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline'])
    print([(k['name'][:14],k['ms'],k.get('frac')) for k in d['kernels']])
    print(d['quality'], d['refit_ms'], d['clocks'], d.get('cpu_baseline'))
print(open("gpurun_out/r1_bench_reference.json").read()[:300])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'folded_seed|gemm_tc_kernel' -s 42 -c 14 -o gpurun_out/r1_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fused_bwd|fused_fwd' -s 30 -c 4 -o gpurun_out/r1_prof_vertex python bench.py --loss-path vertex --steps 3 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -30
timeout 1500 python benchmarks/sweep.py > gpurun_out/r1_sweep.jsonl 2> gpurun_out/r1_sweep.err; tail -2 gpurun_out/r1_sweep.err
compute-sanitizer --tool memcheck --print-limit 20 python benchmarks/sanitize.py 2>&1 | grep -E "COMPUTE-SANITIZER|ERROR SUMMARY|Invalid|sanitizer workload|at 0x|Error" | head -40 > gpurun_out/r1_memcheck.txt; tail -2 gpurun_out/r1_memcheck.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29517 benchmarks/multi_gpu_check.py > /dev/null 2>&1 || echo "multi_gpu_check (1 rank) FAILED"

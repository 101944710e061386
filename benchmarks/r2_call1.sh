# round 2, GPU call 1: new parity tests, full GPU suite, bench line, GEMM probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 -s 2>&1 | grep -E "^\[|MPJPE|refit|gradient rel|tcgen05 3xTF32|passed|failed|Error|error|assert|B=4096|dense|shipped" > gpurun_out/r2_gpu_tests.txt
tail -5 gpurun_out/r2_gpu_tests.txt
timeout 300 python benchmarks/gemm_probe.py > gpurun_out/r2_gemm_probe.jsonl 2> gpurun_out/r2_gemm_probe.err; tail -3 gpurun_out/r2_gemm_probe.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -5 gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench.json"))
for k in ("value","ms_per_step","e2e","roofline","whole_step","quality","refit_ms","secondary","cpu_baseline","gpu_eager_reference","clocks"):
    print(k, json.dumps(d.get(k))[:1500])
print([(k['name'][:18],k['ms'],k.get('frac')) for k in d['kernels']])
print(d['other_loss_path'])
PY
head -c 600 gpurun_out/r2_bench_reference.json

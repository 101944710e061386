mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "gemm or critic_forward" 2>&1 | tail -5
timeout 300 python benchmarks/gemm_probe.py > gpurun_out/r2_gemm_probe2.jsonl 2> gpurun_out/r2_gemm_probe2.err; tail -3 gpurun_out/r2_gemm_probe2.err; cat gpurun_out/r2_gemm_probe2.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s 2>&1 | grep -E "^\[|MPJPE|refit|gradient rel|tcgen05 3xTF32|passed|failed|Error|error|assert|B=4096|dense|shipped" > gpurun_out/r2_gpu_tests.txt
tail -12 gpurun_out/r2_gpu_tests.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -5 gpurun_out/r2_bench_b.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_b.json"))
for k in ("value","ms_per_step","e2e","roofline","whole_step","quality"):
    print(k, json.dumps(d.get(k))[:700])
print([(k['name'][:18],k['ms'],k.get('frac')) for k in d['kernels']])
print(d['other_loss_path'])
PY

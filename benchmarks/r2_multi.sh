mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -4 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
for k in ("value","ms_per_step","n_gpus","e2e","secondary","refit_ms","quality"):
    print(k, json.dumps(d.get(k))[:1200])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 benchmarks/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_${N}.json 2> gpurun_out/r2_mgc.err; tail -2 gpurun_out/r2_mgc.err; head -c 600 gpurun_out/r2_multi_gpu_check_${N}.json

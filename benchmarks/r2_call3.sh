mkdir -p gpurun_out
timeout 300 python benchmarks/tma_probe.py > gpurun_out/r2_tma_probe.jsonl 2> gpurun_out/r2_tma_probe.err; tail -3 gpurun_out/r2_tma_probe.err; cat gpurun_out/r2_tma_probe.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -s -k "bench_config or dense_regressor_100 or shipped_100" 2>&1 | grep -E "^\[|passed|failed|assert" | cut -c1-300

timeout 1500 python benchmarks/sweep.py > gpurun_out/r1_sweep.jsonl 2> gpurun_out/r1_sweep.err; tail -3 gpurun_out/r1_sweep.err; cat gpurun_out/r1_sweep.jsonl | cut -c1-1500

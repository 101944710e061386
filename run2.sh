set -x
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/t_gpu.log; tail -5 gpurun_out/t_gpu.log
python bench.py > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; tail -c 3000 gpurun_out/bench_r1a.json; tail -5 gpurun_out/bench_r1a.err
python bench.py --regressor shipped --no-cpu-baseline > gpurun_out/bench_r1a_shipped.json 2>> gpurun_out/bench_r1a.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_r1a_ref.json 2>> gpurun_out/bench_r1a.err; cat gpurun_out/bench_r1a_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'skin_bwd|skin_fwd|gemm_tc' -s 40 -c 8 -o gpurun_out/prof_r1a python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out

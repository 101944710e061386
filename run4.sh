python -m pytest tests -m gpu -q --timeout 900 -k "refine or smpl_backward" 2>&1 | tail -4
for ov in 0 1; do JRR_OVERLAP_CRITIC=$ov python bench.py --no-cpu-baseline > gpurun_out/bench_ov$ov.json 2> gpurun_out/bench_ov$ov.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_ov$ov.json'))
print("overlap=$ov", d['value'], d['ms_per_step'], d['e2e']['value'], [(k['name'],k['ms']) for k in d['kernels'][:4]])
PY
tail -2 gpurun_out/bench_ov$ov.err; done

timeout 400 python -m pytest tests -m gpu -q --timeout 100 -x -k "single_step or ragged" 2>&1 | tail -6
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
for fb in 0 1; do JRR_FUSED_BWD=$fb timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fused_bwd=$fb', d['value'], d['ms_per_step'], d['e2e']['value'], [(k['name'][:14],k['ms']) for k in d['kernels'][:7]], d['quality'])"; done

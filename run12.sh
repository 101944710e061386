timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], [(k['name'][:14],k['ms']) for k in d['kernels'][:5]], d['quality'])"; done

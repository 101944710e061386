timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6
for r in dense shipped; do timeout 300 python bench.py --no-cpu-baseline --regressor $r 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench $r', d['value'], d['ms_per_step'], d['e2e']['value'], [(k['name'][:14],k['ms']) for k in d['kernels'][:6]], d['quality'], d['refit_ms'])"; done

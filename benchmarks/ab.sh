#!/bin/bash
# usage: ab.sh ENVVAR  -- alternates ENVVAR=1/0 bench runs on one box
for v in 1 0 1 0; do
  env $1=$v python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1=$v', round(d['value']), round(d['ms_per_step'],4), round(d['other_loss_path']['value']))"
done

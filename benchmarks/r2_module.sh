ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_module_launches.csv python benchmarks/module_calls.py > gpurun_out/r2_module.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/r2_module_launches.csv") if not l.startswith("==")]
for row in csv.DictReader(lines):
    v=float(row["Metric Value"].replace(",","")); u=row["Metric Unit"]
    if u=="ns": v/=1000
    elif u=="ms": v*=1000
    print(f'{v:9.2f} us  grid {row["Grid Size"]:>14} blk {row["Block Size"]:>12}  {row["Kernel Name"][:90]}')
PY

# launch list of the module path (fused / unfused kernels): bash benchmarks/r2_module.sh [JRR_FUSED_MODULE value]
F=${1:-1}
JRR_FUSED_MODULE=$F ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_module_launches_f$F.csv python benchmarks/module_calls.py > gpurun_out/r2_module.log 2>&1
python - $F <<'PY'
import csv,sys
lines=[l for l in open(f"gpurun_out/r2_module_launches_f{sys.argv[1]}.csv") if not l.startswith("==")]
rows=list(csv.DictReader(lines))
# last forward+backward of every batch size: print per-kernel time
for row in rows:
    if "at::" in row["Kernel Name"]: 
        print("--")
        continue
    v=float(row["Metric Value"].replace(",","")); u=row["Metric Unit"]
    if u=="ns": v/=1000
    elif u=="ms": v*=1000
    print(f'{v:9.2f} us  grid {row["Grid Size"]:>14}  {row["Kernel Name"][:70]}')
PY

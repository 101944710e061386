timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "find_joints or silhouette or small_batch" 2>&1 | tail -3
bash benchmarks/gpu_validate.sh r2 2
python - <<'PY'
import csv
for name in ("gpurun_out/r2_silhouette_launches.csv",):
    lines=[l for l in open(name) if not l.startswith("==")]
    rows=list(csv.DictReader(lines))
    for row in rows[-45:]:
        if "at::" in row["Kernel Name"]: continue
        v=float(row["Metric Value"].replace(",","")); u=row["Metric Unit"]
        if u=="ns": v/=1000
        elif u=="ms": v*=1000
        print(f'{v:9.2f} us  grid {row["Grid Size"]:>16}  {row["Kernel Name"][:60]}')
PY

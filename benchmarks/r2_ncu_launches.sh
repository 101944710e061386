mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows=[]
with open("gpurun_out/r2_launches.csv") as f:
    lines=[l for l in f if not l.startswith("==")]
r=csv.DictReader(lines)
agg=collections.OrderedDict()
for row in r:
    name=row.get("Kernel Name","")
    try: v=float(row.get("Metric Value","0").replace(",",""))
    except: continue
    unit=row.get("Metric Unit","")
    if unit=="ns": v/=1000.0
    elif unit=="ms": v*=1000.0
    a=agg.setdefault(name,[0,0.0,1e9]); a[0]+=1; a[1]+=v; a[2]=min(a[2],v)
for k,(n,t,mn) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]:
    print(f"{n:4d} x  avg {t/n:8.2f} us  min {mn:8.2f} us   {k[:110]}")
PY

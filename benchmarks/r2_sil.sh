timeout 900 python -m pytest tests/test_silhouette.py -m gpu -x -q --timeout 600 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_silhouette_launches.csv python benchmarks/sil_calls.py > gpurun_out/r2_sil_calls.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/r2_silhouette_launches.csv") if not l.startswith("==")]
rows=list(csv.DictReader(lines))
for row in rows[-31:]:
    if "at::" in row["Kernel Name"]: continue
    v=float(row["Metric Value"].replace(",","")); u=row["Metric Unit"]
    if u=="ns": v/=1000
    elif u=="ms": v*=1000
    print(f'{v:9.2f} us  grid {row["Grid Size"]:>16}  {row["Kernel Name"][:60]}')
PY
python - <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, bench
import jrr_b200 as jrr
dev = torch.device("cuda", 0)
smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
torch.manual_seed(0)
print(json.dumps(bench.run_silhouette(jrr, smpl, torch.rand(17, 6890) + 0.01, jrr.Discriminator().state_dict(), dev)))
print(json.dumps(bench.run_silhouette(jrr, smpl, torch.rand(17, 6890) + 0.01, jrr.Discriminator().state_dict(), dev, n=4096)))
PY

timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -k "folded or fold or refit or loop or six_weight or bench or 100_iterations or find_joints" 2>&1 | tail -4
cat > /tmp/fold_probe.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import jrr_b200 as jrr
dev = torch.device("cuda", 0)
smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
nat = smpl.native(); nat.set_loss_path("folded")
J = torch.rand(17, 6890, device=dev) + 0.01
for _ in range(3): nat.set_regressor(J)
torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_fold_launches.csv python /tmp/fold_probe.py > /dev/null 2>&1
grep -i "fold" gpurun_out/r2_fold_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-100 | tail -5
for cfg in "JRR_FOLD_GEMM=1" "JRR_FOLD_GEMM=0"; do
  env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r2_tmp.json 2> gpurun_out/r2_tmp.err || { echo "$cfg FAILED"; tail -3 gpurun_out/r2_tmp.err; }
  python - "$cfg" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r2_tmp.json").read().strip().splitlines()[-1])
print(sys.argv[1], "value", round(d['value']), d['ms_per_step'], "e2e", round(d['e2e']['value']), d['e2e']['ms_per_step'], d['refit_ms'], d['quality']['mpjpe_after_mm'])
PY
done

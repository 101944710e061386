timeout 900 python -m pytest tests/test_silhouette.py -m gpu -x -q --timeout 600 -s 2>&1 | grep -v "^$" | tail -8 | cut -c1-300
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
grep -i "fold" gpurun_out/r2_fold_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | tail -6

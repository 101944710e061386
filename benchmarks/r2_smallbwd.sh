timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -s -k "small_batch or smpl_backward or six_weight" 2>&1 | grep -v "^$" | tail -16 | cut -c1-250
python - <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
import jrr_b200 as jrr
dev = torch.device("cuda", 0)
smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
pk = {"bf16_burst": 1638.9, "hbm_gbs": 6555.5}
for r in bench.run_c5(jrr, smpl, dev, pk, [1, 8, 9, 16, 32]): print(json.dumps(r))
PY

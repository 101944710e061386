# module path: parity tests, then the C5 sweep rows (forward / forward+backward per batch size)
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -4
python - <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
import jrr_b200 as jrr
dev = torch.device("cuda", 0)
smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
J = torch.rand(17, 6890, device=dev)
smpl.native().set_regressor(J)      # dense regressor, like the bench's secondary section
pk = {"bf16_burst": 1638.9, "hbm_gbs": 6555.5}
rows = bench.run_c5(jrr, smpl, dev, pk, [1, 4, 16, 32, 64, 256, 1024, 4096])
for r in rows: print(json.dumps(r))
PY

# fused forward on CTA pairs: parity tests, then A/B bench of the per-vertex formulation and the C5 sweep
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
for cfg in "JRR_FUSED_PAIR=1" "JRR_FUSED_PAIR=0"; do
  env $cfg timeout 300 python bench.py --loss-path vertex --no-cpu-baseline --no-secondary > gpurun_out/r2_tmp.json 2> gpurun_out/r2_tmp.err || { echo "$cfg FAILED"; tail -3 gpurun_out/r2_tmp.err; }
  python - "$cfg" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r2_tmp.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value']), d['ms_per_step'], round(d['e2e']['value']), [(k['name'][:14],k['ms'],k.get('frac')) for k in d['kernels'][:4]], d['quality']['mpjpe_after_mm'], round(d['other_loss_path']['value']), d['refit_ms'])
PY
done
python - <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
import jrr_b200 as jrr
dev = torch.device("cuda", 0)
smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
J = torch.rand(17, 6890, device=dev)
smpl.native().set_regressor(J)
pk = {"bf16_burst": 1638.9, "hbm_gbs": 6555.5}
rows = bench.run_c5(jrr, smpl, dev, pk, [1, 8, 256, 1024, 4096, 16384])
for r in rows: print(json.dumps(r))
PY

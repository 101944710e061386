# small-batch module forward: parity test, host-side cost per call, kernel times (ncu launch list)
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "small_batch or smpl_forward or smpl_backward or ragged or six_weight" 2>&1 | tail -3
python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import jrr_b200 as jrr
dev = torch.device("cuda", 0)
smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
nat = smpl.native()
base = jrr.synthetic.make_pose_inputs(64, 7)
for B in (1, 4, 8, 16):
    full = torch.from_numpy(base["true_rotmat"])[:B].reshape(B, 24, 9).to(dev).contiguous()
    b = torch.from_numpy(base["true_betas"])[:B].to(dev).contiguous()
    for _ in range(20): nat.smpl_forward(b, full, 0, True, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500): nat.smpl_forward(b, full, 0, True, True)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(50):
        e0.record(); nat.smpl_forward(b, full, 0, True, True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"B={B}: host issue {1e6 * (t1 - t0) / 500:.1f} us/call, back-to-back incl. drain {1e6 * (t2 - t0) / 500:.1f} us/call, event-timed single call median {ts[25]:.1f} us min {ts[0]:.1f}", flush=True)
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_module_launches_f1.csv python benchmarks/module_calls.py > gpurun_out/r2_module.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/r2_module_launches_f1.csv") if not l.startswith("==")]
for row in csv.DictReader(lines):
    if "at::" in row["Kernel Name"]:
        print("--"); continue
    v=float(row["Metric Value"].replace(",","")); u=row["Metric Unit"]
    if u=="ns": v/=1000
    elif u=="ms": v*=1000
    print(f'{v:9.2f} us  grid {row["Grid Size"]:>14}  {row["Kernel Name"][:60]}')
PY

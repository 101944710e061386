mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "smpl or find_joints or transl or full_size" 2>&1 | tail -3
timeout 600 python benchmarks/sweep.py --only c5 --max-log2 12 2>gpurun_out/r2_sweep.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    for r in d.get('rows',[]): print(r)"
tail -2 gpurun_out/r2_sweep.err

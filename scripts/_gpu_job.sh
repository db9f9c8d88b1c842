mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "round_robin" 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python scripts/profile_generic_D.py --dmax 8 2>/dev/null | tee gpurun_out/r3b_generic_D8.json
timeout 900 python scripts/profile_generic_D.py --dmax 16 2>/dev/null | tee gpurun_out/r3b_generic_D16.json
timeout 600 python scripts/run_small_configs.py > gpurun_out/r3b_small.jsonl 2>/dev/null; python -c "
import json
for l in open('gpurun_out/r3b_small.jsonl'):
    d=json.loads(l); print({k:(round(v,6) if isinstance(v,float) else v) for k,v in d.items() if k.startswith('steps_per_s') or 'diff' in k or 'equal' in k})"

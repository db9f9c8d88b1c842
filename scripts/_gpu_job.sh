mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 900 python scripts/profile_generic_D.py --dmax 8 2>/dev/null | tee gpurun_out/r2w_generic_D8.json
timeout 600 python scripts/run_small_configs.py --no-oracle > gpurun_out/r2w_small.jsonl 2>/dev/null; python -c "
import json
for l in open('gpurun_out/r2w_small.jsonl'):
    d=json.loads(l); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k.startswith('steps_per_s') or k.startswith('bloch') or k.startswith('bitstr')})"

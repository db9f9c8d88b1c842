mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python scripts/run_config1_maxcut.py 260 2>/dev/null | tee gpurun_out/r3f_config1_260.json | cut -c100-800
timeout 600 python scripts/run_small_configs.py > gpurun_out/r3f_small.jsonl 2>/dev/null; python -c "
import json
for l in open('gpurun_out/r3f_small.jsonl'):
    d=json.loads(l); print({k:(round(v,6) if isinstance(v,float) else v) for k,v in d.items() if k.startswith('steps_per_s') or 'diff' in k or 'equal' in k})"

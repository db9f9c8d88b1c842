mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 900 python scripts/profile_generic_D.py --dmax 8 2>/dev/null | tee gpurun_out/r2y_generic_D8.json
timeout 1200 python scripts/profile_generic_D.py --dmax 16 2>/dev/null | tee gpurun_out/r2y_generic_D16.json
timeout 600 python scripts/run_small_configs.py --no-oracle > gpurun_out/r2y_small.jsonl 2>/dev/null; python -c "
import json
for l in open('gpurun_out/r2y_small.jsonl'):
    d=json.loads(l); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k.startswith('steps_per_s')})"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r2y_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2y_ncu_launch.log 2>&1; tail -1 gpurun_out/r2y_ncu_launch.log | cut -c1-300

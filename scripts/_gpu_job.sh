mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for z in 0 1; do BQA_B200_EXT_AHEAD=$z python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r2t_bench_ahead$z.json 2>gpurun_out/r2t_bench_ahead$z.err; python -c "
import json; d=json.load(open('gpurun_out/r2t_bench_ahead$z.json')); print('ext ahead $z', d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['kernel_ms_per_step'])" || tail -5 gpurun_out/r2t_bench_ahead$z.err; done
timeout 600 python scripts/run_small_configs.py --no-oracle > gpurun_out/r2t_small.jsonl 2>/dev/null; python -c "
import json
for l in open('gpurun_out/r2t_small.jsonl'):
    d=json.loads(l); print({k:round(v,1) for k,v in d.items() if k.startswith('steps_per_s') and 'one_launch' in k})"

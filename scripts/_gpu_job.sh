mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py 2>gpurun_out/r3i_bench_n1.err | tail -1 > gpurun_out/r3i_bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/r3i_bench_n1.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['parity']['ok'], d['parity']['max_abs'], d['clocks'], d['gpu_launches'], d['kernel_ms_per_step'])"

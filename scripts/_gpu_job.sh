mkdir -p gpurun_out
timeout 600 torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r2v_mg2_check.log 2>&1; tail -4 gpurun_out/r2v_mg2_check.log
timeout 600 python -m pytest tests -m gpu -q -x -k "partitioned or not_current" 2>&1 | tail -3
timeout 900 torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2v_bench_n2.json 2> gpurun_out/r2v_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2v_bench_n2.json')); print(d['value'], d['e2e']['value'], d['parity_vs_1gpu'], d['kernel_ms_per_step'])" || tail -20 gpurun_out/r2v_bench_n2.err
timeout 300 torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/trace_bp_sweeps.py > gpurun_out/r2v_trace_n2.jsonl 2>/dev/null; cat gpurun_out/r2v_trace_n2.jsonl

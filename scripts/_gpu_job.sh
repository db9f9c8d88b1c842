mkdir -p gpurun_out
N=$1
for z in 1 0 1 0; do
BQA_B200_EXT_AHEAD=$z timeout 600 torchrun --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 5 --no-config5 > gpurun_out/r2z_ab_n${N}_ahead$z.json 2> gpurun_out/r2z_ab.err; python -c "
import json; d=json.load(open('gpurun_out/r2z_ab_n${N}_ahead$z.json')); print('ahead $z', round(d['value'],1), round(d['e2e']['value'],1), d['parity_vs_1gpu']['max_abs'], {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()})" || tail -5 gpurun_out/r2z_ab.err
done

#!/bin/bash
# round-2 session 23 (8 GPUs): the final build at N = 8 on the north-star input
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 --workload star4x100M_k25 --no-cpu-baseline --no-wall-clock > gpurun_out/r2s23_bench_star4x100M_k25_n8.json 2> gpurun_out/r2s23_bench_star4x100M_k25_n8.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2s23_bench_star4x100M_k25_n8.json').read().strip().splitlines()[-1])
    print('N=8', 'value %.1fM'%(d['value']/1e6), 'ms/step %.1f'%d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'parity', d['parity_vs_oracle'], 'rounds', d['detail']['rounds'], 'trav ms', d['roofline']['kernel_ms_per_step'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2s23_bench_star4x100M_k25_n8.err').read()[-1500:])
P

#!/bin/bash
# round-2 session 13 (8 GPUs): the peer-mailbox path at N = 4 and 8 (does it run, is it right, what does it give)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2s13_gpus.txt
runb() { # workload N
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2951$2 bench.py --gpus $2 --steps 5 --warmup 3 --workload $1 --no-cpu-baseline --no-wall-clock > gpurun_out/r2s13_bench_$1_n$2.json 2> gpurun_out/r2s13_bench_$1_n$2.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r2s13_bench_$1_n$2.json').read().strip().splitlines()[-1])
    print('$1', 'N=$2', 'value %.1fM'%(d['value']/1e6), 'ms/step %.1f'%d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'parity', d['parity_vs_oracle'], 'rounds', d['detail']['rounds'], 'trav ms', d['roofline']['kernel_ms_per_step'])
except Exception as e:
    print('$1 N=$2 failed', e); print(open('gpurun_out/r2s13_bench_$1_n$2.err').read()[-1500:])
P
}
runb star4x10M_k21 4
runb star4x10M_k21 8
runb star4x100M_k25 8
runb star4x100M_k25 4

#!/bin/bash
# round-2 session 10 (2 GPUs): peer-mailbox multi-GPU path: equality test, bench at N=2 on configs[1] and on the north-star input
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2s10_gpus.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus" > gpurun_out/r2s10_pytest.log 2>&1; tail -5 gpurun_out/r2s10_pytest.log
for wl in star4x10M_k21 star4x100M_k25; do
  timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --workload $wl --no-cpu-baseline --no-wall-clock > gpurun_out/r2s10_bench_${wl}_n1.json 2> gpurun_out/r2s10_bench_${wl}_n1.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload $wl --no-cpu-baseline --no-wall-clock > gpurun_out/r2s10_bench_${wl}_n2.json 2> gpurun_out/r2s10_bench_${wl}_n2.err
  for n in 1 2; do python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r2s10_bench_${wl}_n$n.json').read().strip().splitlines()[-1])
    print('$wl', 'N=$n', 'value %.1fM'%(d['value']/1e6), 'ms/step %.1f'%d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'parity', d['parity_vs_oracle'], 'rounds', d['detail']['rounds'])
except Exception as e:
    print('$wl N=$n failed', e); print(open('gpurun_out/r2s10_bench_${wl}_n$n.err').read()[-1500:])
P
  done
done

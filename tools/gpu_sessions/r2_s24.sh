#!/bin/bash
# round-2 session 24 (last): the Python mirror's single-copy result transfer; parity suite + bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graph.py -m gpu -x -q > gpurun_out/r2s24_pytest.log 2>&1; tail -3 gpurun_out/r2s24_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-wall-clock > gpurun_out/r2s24_bench.json 2> gpurun_out/r2s24_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2s24_bench.json').read().strip().splitlines()[-1])
print('value %.1fM'%(d['value']/1e6),'ms/step %.1f'%d['ms_per_step'],'e2e %.1fM (%.1f ms)'%(d['e2e']['value']/1e6,d['e2e']['ms_per_step']),'parity',d['parity_vs_oracle'],'traffic',d['roofline']['traffic'],'parts',d['detail']['e2e_last_step_parts'])
P
tail -3 gpurun_out/r2s24_bench.err

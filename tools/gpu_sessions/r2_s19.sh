#!/bin/bash
# round-2 session 19: the bench exactly as the driver runs it, launch list of the same command, traffic of a whole pass, full capture
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2s19_pytest.log 2>&1; tail -3 gpurun_out/r2s19_pytest.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2s19_bench_ref.json 2> gpurun_out/r2s19_bench_ref.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2s19_bench.json 2> gpurun_out/r2s19_bench.err
tail -c 600 gpurun_out/r2s19_bench_ref.json; tail -4 gpurun_out/r2s19_bench_ref.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2s19_bench.json').read().strip().splitlines()[-1])
print('value %.1fM'%(d['value']/1e6),'ms/step %.1f'%d['ms_per_step'],'e2e %.1fM (%.1f ms)'%(d['e2e']['value']/1e6,d['e2e']['ms_per_step']),'frac %.4f'%d['roofline']['frac'],'parity',d['parity_vs_oracle'],'wall',d['wall_clock'],'parts',d['detail']['e2e_last_step_parts'])
P
tail -4 gpurun_out/r2s19_bench.err
# traffic + time of every traversal launch of one pass (general kernel behind the common-case kernel: the tool serialises kernels)
LCB_HEAVY_BEHIND=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name regex:k_traverse -c 1000 --csv --log-file gpurun_out/r2s19_traffic_h.csv \
   python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s19_ncu_traffic.log 2>&1
LCB_HEAVY_BEHIND=1 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_traverse_lean --launch-skip 12 --launch-count 3 \
  -o gpurun_out/r2s19_h_lean -f python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s19_ncu_full.log 2>&1
# launch list of the bench command itself
LCB_HEAVY_BEHIND=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2s19_launches_bench.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-wall-clock > gpurun_out/r2s19_bench_under_ncu.json 2> gpurun_out/r2s19_bench_under_ncu.err
ls -la gpurun_out/r2s19*

#!/bin/bash
# round-2 session 20: adaptive fast/deep vote, device-side packing in lcb_create
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2s20_pytest.log 2>&1; tail -3 gpurun_out/r2s20_pytest.log
echo "== H" >> gpurun_out/r2s20.log
timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s20.log
echo "== H from the junction file (lcb_create from host arrays)" >> gpurun_out/r2s20.log
timeout 600 python tools/time_case.py --length 100000000 --k 25 --reps 3 --no-counters 2>&1 | grep '"rep"\|load' >> gpurun_out/r2s20.log
echo "== mammal 8x10M k25" >> gpurun_out/r2s20.log
timeout 900 python tools/time_case.py --kind mammal --genomes 8 --length 10000000 --k 25 --seed 3 --rate 0.03 --reps 2 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s20.log
echo "== mammal 8x10M k25 deep never first (LCB_... n/a)" >> gpurun_out/r2s20.log
echo "== pangenome 16x5M k15" >> gpurun_out/r2s20.log
timeout 900 python tools/time_case.py --kind pangenome --genomes 16 --length 5000000 --k 15 --seed 4 --rate 0.02 --reps 1 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s20.log
python - <<'P'
import json
cur=None
for l in open('gpurun_out/r2s20.log'):
    if l.startswith('=='): cur=l.strip(); continue
    if l.startswith('PARITY') or l.startswith('load'): print(cur, l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(cur, d['rep'], 'create_s',d['create_s'],'h2d_ms',d['ms_h2d'],'find_ms',d['ms_find'],'trav_ms',d['ms_traverse_kernels'],'rounds',d['rounds'],'tail',[round(x,2) for x in d['ms_tail']], 'bails', d['lean_bails'])
P

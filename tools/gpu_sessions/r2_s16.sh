#!/bin/bash
# round-2 session 16: where the rest of a round goes (device-side stamps)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
echo "== H" >> gpurun_out/r2s16.log
timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s16.log
echo "== C2" >> gpurun_out/r2s16.log
timeout 300 python tools/time_case.py --length 10000000 --k 21 --reps 2 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s16.log
python - <<'P'
import json
cur=None
for l in open('gpurun_out/r2s16.log'):
    if l.startswith('=='): cur=l.strip(); continue
    if l.startswith('PARITY'): print(cur, l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(cur, d['rep'], 'find_ms',d['ms_find'],'trav_ms',d['ms_traverse_kernels'],'rounds',d['rounds'],'enum',d['ms_enumerate'],'tail [gap rebase claim diff validate commit]',[round(x,2) for x in d['ms_tail']], 'd2h', d['ms_d2h'])
P

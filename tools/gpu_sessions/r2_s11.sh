#!/bin/bash
# round-2 session 11: cold multi-pass vote out of line; quick check
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s11_pytest.log 2>&1; tail -3 gpurun_out/r2s11_pytest.log
echo "== default H" >> gpurun_out/r2s11_ab.log
timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s11_ab.log
echo "== behind H" >> gpurun_out/r2s11_ab.log
LCB_HEAVY_BEHIND=1 timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s11_ab.log
echo "== default C2" >> gpurun_out/r2s11_ab.log
timeout 300 python tools/time_case.py --length 10000000 --k 21 --reps 3 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s11_ab.log
echo "== mammal 8x10M k25" >> gpurun_out/r2s11_ab.log
timeout 900 python tools/time_case.py --kind mammal --genomes 8 --length 10000000 --k 25 --seed 3 --rate 0.03 --reps 2 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s11_ab.log
LCB_HEAVY_BEHIND=1 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_traverse_lean --launch-skip 70 --launch-count 3 \
  -o gpurun_out/r2s11_h_lean -f python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s11_ncu.log 2>&1
python - <<'P'
import json
cur=None
for l in open('gpurun_out/r2s11_ab.log'):
    if l.startswith('=='): cur=l.strip(); continue
    if l.startswith('PARITY'): print(cur, l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(cur, d['rep'], 'find_ms',d['ms_find'],'trav_ms',d['ms_traverse_kernels'],'rounds',d['rounds'],'lean',d.get('lean_runs'),d.get('lean_bails'),'why',d.get('lean_bail_why'))
P

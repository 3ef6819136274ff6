#!/bin/bash
# round-2 session 15: tail of a round with 2.5-ms rounds (launch list), lane-per-seed k_claim
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2s15_pytest.log 2>&1; tail -3 gpurun_out/r2s15_pytest.log
echo "== default H" >> gpurun_out/r2s15_ab.log
timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s15_ab.log
echo "== default C2" >> gpurun_out/r2s15_ab.log
timeout 300 python tools/time_case.py --length 10000000 --k 21 --reps 2 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s15_ab.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2s15_launches_h.csv \
   python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s15_ncu_launches.log 2>&1
python - <<'P'
import json, csv, collections
cur=None
for l in open('gpurun_out/r2s15_ab.log'):
    if l.startswith('=='): cur=l.strip(); continue
    if l.startswith('PARITY'): print(cur, l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(cur, d['rep'], 'find_ms',d['ms_find'],'trav_ms',d['ms_traverse_kernels'],'rounds',d['rounds'],'enum',d['ms_enumerate'])
rows=[r for r in csv.reader(open('gpurun_out/r2s15_launches_h.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[hdr+1:]:
    try: v=float(r[mv].replace(',',''))
    except: continue
    n=r[kn].split('(')[0][:40]; tot[n]+=v; cnt[n]+=1
for n,v in tot.most_common(24): print("%-42s %6d launches %10.3f ms total %8.1f us avg"%(n,cnt[n],v/1e6, v/1e3/cnt[n]))
P

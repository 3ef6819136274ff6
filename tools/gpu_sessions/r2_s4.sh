#!/bin/bash
# round-2 session 4: launch list of one H pass (where does the non-traversal time go), wave quantisation, micro-optimised lean kernel
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() {
  label=$1; shift
  echo "== $label H" >> gpurun_out/r2s4_ab.log
  env "$@" timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s4_ab.log
  echo "== $label C2" >> gpurun_out/r2s4_ab.log
  env "$@" timeout 300 python tools/time_case.py --length 10000000 --k 21 --reps 3 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s4_ab.log
}
run default X=1
run wave LCB_WAVE_QUANT=1
run hostloop LCB_HOST_LOOP=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2s4_launches_h.csv \
   python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s4_ncu_launches.log 2>&1
python - <<'P'
import json, csv, collections
cur=None
for l in open('gpurun_out/r2s4_ab.log'):
    if l.startswith('=='): cur=l.strip(); continue
    if l.startswith('PARITY'): print(cur, l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(cur, d['rep'], 'find_ms',d['ms_find'],'trav_ms',d['ms_traverse_kernels'],'rounds',d['rounds'],'why',d.get('lean_bail_why'))
rows=[r for r in csv.reader(open('gpurun_out/r2s4_launches_h.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[hdr+1:]:
    try: v=float(r[mv].replace(',',''))
    except: continue
    n=r[kn].split('(')[0][:40]; tot[n]+=v; cnt[n]+=1
S=sum(tot.values())
for n,v in tot.most_common(25): print("%-42s %6d launches %10.1f us %5.1f%%"%(n,cnt[n],v/1000.0 if False else v, 100*v/S))
P

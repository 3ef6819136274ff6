#!/bin/bash
# round-2 session 12: smaller hot code (two-tier fast vote, cold helpers out of line); round-length sweep; launch list; CLI timeline
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s12_pytest.log 2>&1; tail -3 gpurun_out/r2s12_pytest.log
run() {
  label=$1; shift
  echo "== $label H" >> gpurun_out/r2s12_ab.log
  env "$@" timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 2 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s12_ab.log
  echo "== $label C2" >> gpurun_out/r2s12_ab.log
  env "$@" timeout 300 python tools/time_case.py --length 10000000 --k 21 --reps 2 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s12_ab.log
}
run default X=1
run minround1.0 LCB_MIN_ROUND_MS=1.0
run minround1.5 LCB_MIN_ROUND_MS=1.5
run minround2.5 LCB_MIN_ROUND_MS=2.5
run grow1.6 LCB_GROW_BELOW=1.6 LCB_SHRINK_ABOVE=3.0
echo "== mammal 8x10M k25" >> gpurun_out/r2s12_ab.log
timeout 900 python tools/time_case.py --kind mammal --genomes 8 --length 10000000 --k 25 --seed 3 --rate 0.03 --reps 2 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s12_ab.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2s12_launches_h.csv \
   python tools/time_case.py --length 100000000 --k 25 --reps 1 --construct --no-counters > gpurun_out/r2s12_ncu_launches.log 2>&1
# CLI timeline on the junction file of the bench directory (made by the reference twopaco)
python - <<'P' > gpurun_out/r2s12_prep.log 2>&1
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import bench
print(bench.prepare_workload('star4x100M_k25', 0))
P
timeout 600 python tools/time_cli_h.py /tmp/sibeliaz_b200_bench/star4x100M_k25/g.dbg > gpurun_out/r2s12_cli_trace.log 2>&1
python - <<'P'
import json, csv, collections
cur=None
for l in open('gpurun_out/r2s12_ab.log'):
    if l.startswith('=='): cur=l.strip(); continue
    if l.startswith('PARITY'): print(cur, l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(cur, d['rep'], 'find_ms',d['ms_find'],'trav_ms',d['ms_traverse_kernels'],'rounds',d['rounds'],'evals',d['traversals_first']+d['traversals_rerun'],'bails',d.get('lean_bails'))
rows=[r for r in csv.reader(open('gpurun_out/r2s12_launches_h.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[hdr+1:]:
    try: v=float(r[mv].replace(',',''))
    except: continue
    n=r[kn].split('(')[0][:40]; tot[n]+=v; cnt[n]+=1
for n,v in tot.most_common(16): print("%-42s %6d launches %10.3f ms total %8.1f us avg"%(n,cnt[n],v/1e6, v/1e3/cnt[n]))
P
tail -40 gpurun_out/r2s12_cli_trace.log

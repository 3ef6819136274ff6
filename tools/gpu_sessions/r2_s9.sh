#!/bin/bash
# round-2 session 9: two-body vote, long paths in the common-case kernel; the bench exactly as the driver runs it
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s9_pytest.log 2>&1; tail -3 gpurun_out/r2s9_pytest.log
echo "== default H" >> gpurun_out/r2s9_ab.log
timeout 300 python tools/time_case.py --length 100000000 --k 25 --reps 3 --construct --no-counters 2>&1 | grep '"rep"' >> gpurun_out/r2s9_ab.log
echo "== default C2" >> gpurun_out/r2s9_ab.log
timeout 300 python tools/time_case.py --length 10000000 --k 21 --reps 3 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s9_ab.log
echo "== pangenome 16x5M k15" >> gpurun_out/r2s9_ab.log
timeout 900 python tools/time_case.py --kind pangenome --genomes 16 --length 5000000 --k 15 --seed 4 --rate 0.02 --reps 2 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s9_ab.log
echo "== mammal 8x10M k25" >> gpurun_out/r2s9_ab.log
timeout 900 python tools/time_case.py --kind mammal --genomes 8 --length 10000000 --k 25 --seed 3 --rate 0.03 --reps 2 --construct --no-counters --oracle 2>&1 | grep '"rep"\|PARITY' >> gpurun_out/r2s9_ab.log
python - <<'P'
import json
cur=None
for l in open('gpurun_out/r2s9_ab.log'):
    if l.startswith('=='): cur=l.strip(); continue
    if l.startswith('PARITY'): print(cur, l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(cur, d['rep'], 'find_ms',d['ms_find'],'trav_ms',d['ms_traverse_kernels'],'rounds',d['rounds'],'lean',d.get('lean_runs'),d.get('lean_bails'),'why',d.get('lean_bail_why'))
P
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2s9_bench_ref.json 2> gpurun_out/r2s9_bench_ref.err
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2s9_bench.json 2> gpurun_out/r2s9_bench.err
tail -c 1500 gpurun_out/r2s9_bench_ref.json; tail -4 gpurun_out/r2s9_bench_ref.err
tail -c 3500 gpurun_out/r2s9_bench.json; tail -4 gpurun_out/r2s9_bench.err

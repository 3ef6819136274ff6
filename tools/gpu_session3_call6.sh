#!/bin/bash
# Developer helper: two-tier push (straight-line fast path + grouped fallback) against the old serial fallback;
# GPU suite, bench, ncu launch list and full capture of k_traverse for the final build of the session.
mkdir -p gpurun_out
O=gpurun_out/r1s3f
L=$PWD/sibeliaz_b200/lib
python tools/time_case.py --construct --reps 1 > /dev/null 2>&1
for v in "" _pushold; do
  LCB_LIB_PATH=$L/libsibeliaz_lcb$v.so timeout 300 python tools/time_case.py --construct --reps 3 > ${O}_c2$v.log 2>&1; echo "c2 $v rc=$?"
  grep -o '"find_s": [0-9.]*\|"ms_traverse_kernels": [0-9.]*\|"rounds": [0-9]*' ${O}_c2$v.log | paste - - - | tail -2
done
timeout 300 python tools/time_case.py --construct --length 100000000 --k 25 --reps 2 > ${O}_h.log 2>&1; echo "headline rc=$?"
grep -o '"find_s": [0-9.]*\|"ms_traverse_kernels": [0-9.]*\|"rounds": [0-9]*' ${O}_h.log | paste - - - | tail -1
timeout 300 python tools/time_case.py --kind pangenome --genomes 16 --length 5000000 --k 15 --rate 0.02 --seed 4 --construct --oracle --reps 1 > ${O}_pangenome.log 2>&1; echo "pangenome rc=$?"; tail -3 ${O}_pangenome.log | cut -c1-330
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 ${O}_pytest.log
timeout 300 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; cut -c1-330 ${O}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_bench_under_ncu.json 2> ${O}_bench_under_ncu.err; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_traverse -s 1 -c 4 -f -o gpurun_out/prof_traverse_r1s3 python tools/time_case.py --construct --reps 1 > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep

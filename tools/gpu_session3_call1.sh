#!/bin/bash
# Developer helper (one gpurun call): GPU test-suite, fused pipeline on the bench workload / the headline input,
# and the BASELINE configs[2]/[3] kinds against the oracle.  Logs go to gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out/r1s3
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
timeout 300 python tools/time_case.py --construct --reps 2 --oracle > ${O}_c2_construct.log 2>&1; echo "c2 construct rc=$?"; tail -4 ${O}_c2_construct.log | cut -c1-600
timeout 400 python tools/time_case.py --kind mammal --genomes 8 --length 10000000 --k 25 --rate 0.03 --seed 3 --construct --oracle --reps 2 > ${O}_mammal8x10M.log 2>&1; echo "mammal rc=$?"; tail -4 ${O}_mammal8x10M.log | cut -c1-600
timeout 600 python tools/time_case.py --kind pangenome --genomes 16 --length 5000000 --k 15 --rate 0.02 --seed 4 --construct --oracle --reps 2 > ${O}_pangenome16x5M.log 2>&1; echo "pangenome rc=$?"; tail -4 ${O}_pangenome16x5M.log | cut -c1-600
timeout 400 python tools/time_case.py --construct --length 100000000 --k 25 --reps 2 > ${O}_h_construct.log 2>&1; echo "headline construct rc=$?"; tail -3 ${O}_h_construct.log | cut -c1-600
timeout 300 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; cut -c1-900 ${O}_bench.json

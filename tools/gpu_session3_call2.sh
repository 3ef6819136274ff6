#!/bin/bash
# Developer helper (one gpurun call): big-arena re-run path (tiny-arena test build + the 16 x 5 Mbp k=15 pan-genome of
# BASELINE configs[3]), the whole GPU suite, bench.
mkdir -p gpurun_out
O=gpurun_out/r1s3b
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "big_arena" > ${O}_bigarena.log 2>&1; echo "big arena rc=$?"; tail -15 ${O}_bigarena.log
timeout 900 python tools/time_case.py --kind pangenome --genomes 16 --length 5000000 --k 15 --rate 0.02 --seed 4 --construct --oracle --reps 2 > ${O}_pangenome16x5M.log 2>&1; echo "pangenome rc=$?"; tail -4 ${O}_pangenome16x5M.log | cut -c1-900
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
timeout 300 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; cut -c1-400 ${O}_bench.json

#!/bin/bash
# Developer helper: grouped parallel push + bisection + heavy-seed admission: GPU suite (normal and self-check builds),
# the 16 x 5 Mbp k=15 pan-genome with round trace and oracle parity, bench.
mkdir -p gpurun_out
O=gpurun_out/r1s3d
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
LCB_LIB_PATH=$PWD/sibeliaz_b200/lib/libsibeliaz_lcb_check.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not big_arena and not two_gpus" > ${O}_pytest_selfcheck.log 2>&1; echo "selfcheck pytest rc=$?"; tail -3 ${O}_pytest_selfcheck.log
LCB_TRACE_ROUNDS=1 timeout 600 python tools/time_case.py --kind pangenome --genomes 16 --length 5000000 --k 15 --rate 0.02 --seed 4 --construct --oracle --reps 1 > ${O}_pangenome_trace.log 2>&1; echo "pangenome rc=$?"; grep "\[round\]" ${O}_pangenome_trace.log | cut -c1-200 | head -30; tail -3 ${O}_pangenome_trace.log | cut -c1-500
timeout 300 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; cut -c1-330 ${O}_bench.json
timeout 300 python tools/time_case.py --construct --length 100000000 --k 25 --reps 2 > ${O}_h_construct.log 2>&1; echo "headline construct rc=$?"; tail -2 ${O}_h_construct.log | cut -c1-300

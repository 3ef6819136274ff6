#!/bin/bash
# Developer helper: A/B of the push / search variants on the bench workload and the headline input
mkdir -p gpurun_out
O=gpurun_out/r1s3e
L=$PWD/sibeliaz_b200/lib
python tools/time_case.py --construct --reps 1 > /dev/null 2>&1   # generate + warm the page cache
for v in "" _pushold _ubold _pushold_ubold; do
  LCB_LIB_PATH=$L/libsibeliaz_lcb$v.so timeout 300 python tools/time_case.py --construct --reps 3 > ${O}_c2$v.log 2>&1; echo "c2 $v rc=$?"
  grep -o '"find_s": [0-9.]*\|"ms_traverse_kernels": [0-9.]*\|"rounds": [0-9]*' ${O}_c2$v.log | paste - - - | tail -2
done
for v in "" _pushold; do
  LCB_LIB_PATH=$L/libsibeliaz_lcb$v.so timeout 300 python tools/time_case.py --construct --length 100000000 --k 25 --reps 2 > ${O}_h$v.log 2>&1; echo "headline $v rc=$?"
  grep -o '"find_s": [0-9.]*\|"ms_traverse_kernels": [0-9.]*\|"rounds": [0-9]*' ${O}_h$v.log | paste - - - | tail -1
done

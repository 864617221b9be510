#!/bin/bash
set -u
out=gpurun_out/r2_call2
mkdir -p $out
timeout 400 scripts/microbench/gather_bench 10000000 17 3.94 5 > $out/gather_bench_hub.txt 2>&1
timeout 400 scripts/microbench/gather_bench 10000000 17 1.0 3 > $out/gather_bench_uniform.txt 2>&1
timeout 400 scripts/microbench/gather_bench 1000000 17 1.0 5 > $out/gather_bench_l2resident.txt 2>&1
cat $out/gather_bench_hub.txt

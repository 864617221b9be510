#!/bin/bash
# single-GPU call: ncu evidence of the round (launch list of the iteration, --set full of the gather and Lanczos kernels), low-rank roofline
set -u
out=gpurun_out/r2_ncu
mkdir -p $out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-solve --lanczos 5"
timeout 300 python scripts/lowrank_roofline.py > $out/lowrank_roofline.json 2> $out/lowrank_roofline.err; tail -2 $out/lowrank_roofline.json
$B > $out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_rows|k_gram|k_A_rowc|k_step_grad|k_tail_rest|k_biquadratic|k_lz_|k_obj_slots" -c 200 \
    --csv --log-file $out/launches.csv $B > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_rows_group|k_rows_warp" -s 3 -c 3 -o $out/gather_full $B > $out/ncu_gather.log 2>&1
echo "gather full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_lz_spmv" -s 6 -c 3 -o $out/lanczos_full $B > $out/ncu_lz.log 2>&1
echo "lanczos full rc=$?"
ls -la $out

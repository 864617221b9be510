#!/bin/bash
# First GPU call of round 2: the options written blind at the end of round 1 (no GPU minutes were left).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/r2_first_call.sh'
# 1. parity of the experimental kernel configurations (tests/conftest.py EXPERIMENTAL_CONFIGS) on the small families
# 2. C5 bench, default against spmm_prefetch (8- and 4-nonzero blocks); the last_iterate fields must agree with the
#    default's to ~1e-12 (same summation order), the "spmm" section is the number to read
# 3. if (2) is green: per-launch list of the prefetch run for profiles/
# Everything lands in gpurun_out/r2_first/.
set -u
out=gpurun_out/r2_first
mkdir -p $out
export SDPLRP_TEST_EXPERIMENTAL=1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "prefetch or bundle or batched or preprocess_device" > $out/pytest_experimental.log 2>&1
echo "pytest experimental rc=$?" | tee $out/rc.txt
unset SDPLRP_TEST_EXPERIMENTAL

line() {  # name, extra bench args
  name=$1; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "it/s", round(d["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()),
          "L=%.15g obj=%.15g alpha=%.15g" % (d["last_iterate"]["L"], d["last_iterate"]["obj"], d["last_iterate"]["alpha"]),
          ("lanczos_ms_per_step=%.4f" % d["lanczos"]["ms_per_step"]) if d.get("lanczos") else "",
          "setup=%s" % d.get("setup"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
{
line default
line prefetch8 --option spmm_prefetch=1
line prefetch4 --option spmm_prefetch=1 --option spmm_unroll=4
line prefetch8_phases --option spmm_prefetch=1 --option spmm_phases=1
line prefetch8_pad --option spmm_prefetch=1 --option spmm_pad=1
line prefetch4_pad --option spmm_prefetch=1 --option spmm_unroll=4 --option spmm_pad=1
line batched8 --option spmm_prefetch=3
line batched4 --option spmm_prefetch=3 --option spmm_unroll=4
line bundle8 --option spmm_prefetch=2
line bundle4 --option spmm_prefetch=2 --option spmm_unroll=4
line bundle8_pad --option spmm_prefetch=2 --option spmm_pad=1
line bundle4_pad --option spmm_prefetch=2 --option spmm_unroll=4 --option spmm_pad=1
# 8 lanes per row (a lane group = one quarter-warp phase of an LDG.128) on 128-byte rows: one line per phase
line bundle8_pad_g8 --option spmm_prefetch=2 --option spmm_pad=1 --option spmm_g0=0
line prefetch8_pad_g8 --option spmm_prefetch=1 --option spmm_pad=1 --option spmm_g0=0
# hypothesis test for the wavefront model (profiles/r1_gather_size_sweep.md): lines touched per gathered row.
# rank 8 = 64-byte rows (never cross a 128-byte line), 10 = 80 bytes (cross 5 times out of 8), 12 = 96 (6 of 8), 16 = 128 (never)
line rank8 --rank 8
line rank12 --rank 12
line rank16 --rank 16
line g0_pow2 --option spmm_g0=0
line lanczos_default --lanczos 50
line lanczos_bundle --lanczos 50 --option lanczos_bundle=1
line device_triplets --device-triplets
} | tee $out/summary.txt

# stall reasons / L1 wavefronts of the default class-0 kernel (one launch, --set full; read with ncu -i ... --page raw --csv)
ncu --set full --import-source on --clock-control none -k regex:k_rows_group -s 2 -c 1 -o $out/rows_group_default \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_full_default.log 2>&1
if grep -q "^prefetch8 it/s" $out/summary.txt; then
  ncu --set full --import-source on --clock-control none -k regex:k_rows_group_pf -s 2 -c 1 -o $out/rows_group_prefetch8 \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --option spmm_prefetch=1 > $out/ncu_full_prefetch8.log 2>&1
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:k_rows -c 24 --csv --log-file $out/ncu_rows_prefetch8.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --option spmm_prefetch=1 > $out/ncu_prefetch8.log 2>&1
fi
echo done

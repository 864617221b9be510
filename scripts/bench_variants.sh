#!/bin/bash
# A/B of the SpMM knobs on the C5 workload (device-resident section timers only)
for cfg in "SDPLRP_SPMM_UNROLL=8 SDPLRP_SPMM_G0=1" "SDPLRP_SPMM_UNROLL=4 SDPLRP_SPMM_G0=1" "SDPLRP_SPMM_UNROLL=8 SDPLRP_SPMM_G0=0" "SDPLRP_SPMM_UNROLL=4 SDPLRP_SPMM_G0=0 SDPLRP_HOT_ROWS=0" "SDPLRP_SPMM_UNROLL=8 SDPLRP_SPMM_G0=1 SDPLRP_HOT_ROWS=0" "SDPLRP_SPMM_UNROLL=8 SDPLRP_SPMM_G0=1 SDPLRP_HOT_ROWS=1200000"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('it/s', round(d['value'],2), ' '.join(f\"{k}={v['ms_per_iter']:.2f}\" for k,v in d['roofline']['kernels'].items()))"
done

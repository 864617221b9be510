#!/bin/bash
# gather-pass cost versus row size (rank 8 = 64-byte rows, 10 = 80, 12 = 96, 16 = 128)
for r in 8 10 12 16; do
  echo "== rank $r"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --rank $r 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('it/s', round(d['value'],2), ' '.join(f\"{k}={v['ms_per_iter']:.2f}\" for k,v in d['roofline']['kernels'].items()))"
done

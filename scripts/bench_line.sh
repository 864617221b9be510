#!/bin/bash
# one short bench run, sections on one line
timeout 300 python bench.py --steps ${1:-10} --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('it/s', round(d['value'],2), ' '.join(f\"{k}={v['ms_per_iter']:.2f}\" for k,v in d['roofline']['kernels'].items()))"

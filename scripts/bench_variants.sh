#!/bin/bash
# A/B of the L2 knobs of the gather pass on the C5 workload (device-resident section timers only)
for cfg in "SDPLRP_L2_PERSIST_MB=100000" "SDPLRP_L2_PERSIST_MB=0" "SDPLRP_L2_PERSIST_MB=100000 SDPLRP_HOT_ROWS=300000" "SDPLRP_L2_PERSIST_MB=100000 SDPLRP_HOT_ROWS=600000" "SDPLRP_L2_PERSIST_MB=100000 SDPLRP_HOT_ROWS=1500000" "SDPLRP_L2_PERSIST_MB=32"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('it/s', round(d['value'],2), ' '.join(f\"{k}={v['ms_per_iter']:.2f}\" for k,v in d['roofline']['kernels'].items()))"
done
python -c "
import torch
p=torch.cuda.get_device_properties(0); print('L2', p.L2_cache_size)
import ctypes; rt=ctypes.CDLL('libcudart.so.12'); v=ctypes.c_int(); rt.cudaDeviceGetAttribute(ctypes.byref(v), 108, 0); print('maxPersistingL2', v.value)
"

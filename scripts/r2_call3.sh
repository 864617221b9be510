#!/bin/bash
set -u
out=gpurun_out/r2_call3
mkdir -p $out
timeout 300 scripts/microbench/l2_probe > $out/l2_probe_default.txt 2>&1
timeout 300 scripts/microbench/l2_probe 32 > $out/l2_probe_fetch32.txt 2>&1
cat $out/l2_probe_default.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gather" > $out/pytest_gather.log 2>&1
echo "pytest gather rc=$?" | tee $out/rc.txt
tail -15 $out/pytest_gather.log
line() {
  name=$1; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "it/s", round(d["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()),
          "L=%.15g obj=%.15g alpha=%.15g" % (d["last_iterate"]["L"], d["last_iterate"]["obj"], d["last_iterate"]["alpha"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
{
line gather_async --option gather_mode=2
line gather_async_t64 --option gather_mode=2 --option gather_tile=64
line gather_async_s3 --option gather_mode=2 --option gather_tile=64 --option gather_stages=3
line gather_bulk --option gather_mode=1 --option gather_tile=64
} | tee $out/summary.txt
echo done

#!/bin/bash
set -u
out=gpurun_out/r2_call5
mkdir -p $out
line() {
  name=$1; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "it/s", round(d["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()),
          "L=%.15g obj=%.15g" % (d["last_iterate"]["L"], d["last_iterate"]["obj"]),
          ("lanczos_ms_per_step=%.4f" % d["lanczos"]["ms_per_step"]) if d.get("lanczos") else "")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
{
line win32 --option l2_window_mb=32
line win24 --option l2_window_mb=24
line win40 --option l2_window_mb=40
line win16 --option l2_window_mb=16
} | tee $out/summary.txt
( time timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "c5_against or cr_recurrence or preprocess_device" ) > $out/pytest_new.log 2>&1
echo "pytest new rc=$?" | tee $out/rc.txt
tail -12 $out/pytest_new.log
echo done

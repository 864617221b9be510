#!/bin/bash
set -u
out=gpurun_out/r2_call4
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gather" > $out/pytest_gather.log 2>&1
echo "pytest gather rc=$?" | tee $out/rc.txt
tail -8 $out/pytest_gather.log
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
line default --lanczos 50
line win32 --lanczos 50 --option l2_window_mb=32
line win16 --lanczos 50 --option l2_window_mb=16
line win48 --lanczos 50 --option l2_window_mb=48
line win79 --lanczos 50 --option l2_window_mb=79
line win32_reset --lanczos 50 --option l2_window_mb=32 --option l2_window_reset=1
line gather_async --option gather_mode=2
line gather_async_win32 --option gather_mode=2 --option l2_window_mb=32
line gather_bulk_t64 --option gather_mode=1 --option gather_tile=64
} | tee $out/summary.txt
echo done

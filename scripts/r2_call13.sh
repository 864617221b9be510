#!/bin/bash
# A/B (historical: the option "l2_window_mb" and the WIN kernels it selected were removed after this call, commit b70cf7f has them): access-policy window over the hub prefix of the gathered factor in the objective pass
set -u
out=gpurun_out/r2_call13
mkdir -p $out
( time timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "window and (rank_sweep or f_g_linesearch or free_running or step_g)" ) > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt; tail -3 $out/pytest.log
line() {
  name=$1; shift
  timeout 120 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-solve --lanczos 0 "$@" > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "it/s", round(d["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()),
          "L=%.15g" % d["last_iterate"]["L"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
{
line win32 --option l2_window_mb=32
line win48 --option l2_window_mb=48
line win20 --option l2_window_mb=20
line default
} | tee $out/summary.txt

#!/bin/bash
# grid sizes of the streaming kernels: L-BFGS passes ("lb_ctas"), per-row constraint pass ("rowc_ctas"); the tail is at its measured optimum (auto = 6)
set -u
out=gpurun_out/r2_call12
mkdir -p $out
( time timeout 420 python -m pytest tests/test_gpu_driver.py tests/test_gpu_parity.py -q -m gpu -x -k "default or relabel" ) > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt; tail -4 $out/pytest.log
line() {
  name=$1; shift
  timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-solve --lanczos 0 "$@" > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "it/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}({b['frac']:.2f})" for a, b in k.items()),
          "L=%.15g" % d["last_iterate"]["L"], "launches", d["gpu_launches"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
{
line default
line lb3_rowc4 --option lb_ctas=3 --option rowc_ctas=4
line lb5_rowc6 --option lb_ctas=5 --option rowc_ctas=6
line lb6_rowc8 --option lb_ctas=6 --option rowc_ctas=8
line lb7_rowc12 --option lb_ctas=7 --option rowc_ctas=12
line lb8_rowc24 --option lb_ctas=8 --option rowc_ctas=24
} | tee $out/summary.txt
ls $out

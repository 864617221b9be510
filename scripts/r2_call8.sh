#!/bin/bash
set -u
out=gpurun_out/r2_call8
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "default or relabel or phases" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt; tail -3 $out/pytest.log
line() {
  name=$1; shift
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-solve "$@" > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "it/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()),
          "L=%.15g" % d["last_iterate"]["L"], "lanczos", d["lanczos"]["ms_per_step"] if d.get("lanczos") else None, "setup", d["setup"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
{
line default
line gmax64 --option row_group_max=64
line gmax48 --option row_group_max=48
line gmax24 --option row_group_max=24
} | tee $out/summary.txt

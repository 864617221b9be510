#!/bin/bash
set -u
out=gpurun_out/r2_call7
mkdir -p $out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-solve > $out/bench_default.json 2> $out/bench_default.err
python - $out/bench_default.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("default it/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()), "lanczos", d["lanczos"], "setup", d["setup"])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "default or phases" > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log

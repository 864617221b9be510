#!/bin/bash
set -u
out=gpurun_out/r2_call6
mkdir -p $out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > $out/pytest_gpu.log 2>&1
echo "pytest gpu rc=$?" | tee $out/rc.txt
tail -6 $out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-solve > $out/bench_default.json 2> $out/bench_default.err
python - $out/bench_default.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("default it/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()), "lanczos", d["lanczos"], "setup", d["setup"])
PY
SDPLRP_CLASS_STREAMS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-solve > $out/bench_nostreams.json 2> $out/bench_nostreams.err
python - $out/bench_nostreams.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("no class streams it/s", round(d["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()))
PY
bash scripts/r2_ncu.sh

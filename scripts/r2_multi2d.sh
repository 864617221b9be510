#!/bin/bash
set -u
N=2
out=gpurun_out/r2_multi_n2d
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "2" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt; tail -3 $out/pytest.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_auto.json 2> $out/bench_auto.err
python - $out/bench_auto.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("auto gpus", d["n_gpus"], "it/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), "comm_ms", round(d["comm_ms_per_step"], 3),
      " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()), "lanczos", d["lanczos"], "ttt", d["time_to_tol"], "cpu", d["cpu_baseline"])
PY
tail -3 $out/bench_auto.err

#!/bin/bash
set -u
N=2
out=gpurun_out/r2_multi_n2c
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "2" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt; tail -3 $out/pytest.log
timeout 300 python scripts/check_multigpu.py 2 > $out/check_1v2.log 2>&1; echo "check rc=$?" | tee -a $out/rc.txt; tail -8 $out/check_1v2.log
show() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "gpus", d["n_gpus"], "it/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), "comm_ms", round(d["comm_ms_per_step"], 3),
          " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()), "L=%.15g" % d["last_iterate"]["L"], "lanczos", d["lanczos"]["ms_per_step"] if d.get("lanczos") else None)
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run() {
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-solve "$@" > $out/bench_$name.json 2> $out/bench_$name.err
  show $out/bench_$name.json $name | tee -a $out/summary.txt
}
run halo1
run halo2 --option halo=2
echo done

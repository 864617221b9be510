#!/bin/bash
# 8-GPU box: two-phase (halo=1) against single-sweep (halo=2) halo exchange at 8 and 4 GPUs; time-to-tol at 8
set -u
out=gpurun_out/r2_multi8c
mkdir -p $out
show() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "gpus", d["n_gpus"], "it/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), "comm_ms", round(d["comm_ms_per_step"], 3),
          " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()), "L=%.15g obj=%.15g" % (d["last_iterate"]["L"], d["last_iterate"]["obj"]),
          "lanczos", d["lanczos"]["ms_per_step"] if d.get("lanczos") else None, "setup", d.get("setup"), "ttt", d.get("time_to_tol"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run() {
  N=$1; name=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline "$@" > $out/bench_${name}_n$N.json 2> $out/bench_${name}_n$N.err
  show $out/bench_${name}_n$N.json ${name}_n$N | tee -a $out/summary.txt
}
run 8 halo1
run 8 halo2 --no-solve --option halo=2
run 4 halo1 --no-solve
run 4 halo2 --no-solve --option halo=2
echo done

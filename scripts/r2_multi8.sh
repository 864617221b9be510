#!/bin/bash
# 8-GPU call: C5 bench with the halo exchange (default) incl. time-to-tol, and the round-1 all-gather path for comparison
set -u
out=gpurun_out/r2_multi8
mkdir -p $out
N=${1:-8}
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export -f runN; export N
show() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "gpus", d["n_gpus"], "it/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), "comm_ms", round(d["comm_ms_per_step"], 3),
          " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()), "L=%.15g obj=%.15g" % (d["last_iterate"]["L"], d["last_iterate"]["obj"]),
          "lanczos", d.get("lanczos"), "setup", d.get("setup"), "ttt", d.get("time_to_tol"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
timeout 600 bash -c "runN 29581 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline" > $out/bench_halo1_n$N.json 2> $out/bench_halo1_n$N.err
show $out/bench_halo1_n$N.json halo1 | tee $out/summary_n$N.txt
tail -3 $out/bench_halo1_n$N.err
timeout 300 bash -c "runN 29582 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-solve --option halo=0" > $out/bench_halo0_n$N.json 2> $out/bench_halo0_n$N.err
show $out/bench_halo0_n$N.json halo0 | tee -a $out/summary_n$N.txt
echo done

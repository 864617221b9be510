#!/bin/bash
# 8-GPU box, final: the default configuration (halo auto) at 8 GPUs with time-to-tol, and at 4 GPUs
set -u
out=gpurun_out/r2_multi8d
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
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 30 --warmup 5 > $out/bench_n8.json 2> $out/bench_n8.err
show $out/bench_n8.json auto_n8 | tee -a $out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 4 --steps 30 --warmup 5 --no-solve > $out/bench_n4.json 2> $out/bench_n4.err
show $out/bench_n4.json auto_n4 | tee -a $out/summary.txt
echo done

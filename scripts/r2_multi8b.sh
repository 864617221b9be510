#!/bin/bash
# 8-GPU box: multi-GPU parity tests on 2 / 4 / 8 ranks, then the C5 bench at 8 (with time-to-tol), 4 and 2 GPUs
set -u
out=gpurun_out/r2_multi8b
mkdir -p $out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu ) > $out/pytest_multi.log 2>&1
echo "pytest multi rc=$?" | tee $out/rc.txt
tail -12 $out/pytest_multi.log
show() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "gpus", d["n_gpus"], "it/s", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), "comm_ms", round(d["comm_ms_per_step"], 3),
          " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()), "L=%.15g obj=%.15g" % (d["last_iterate"]["L"], d["last_iterate"]["obj"]),
          "lanczos", d.get("lanczos"), "setup", d.get("setup"), "halo", d.get("halo"), "ttt", d.get("time_to_tol"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for N in 8 4 2; do
  extra="--no-solve"; [ $N -eq 8 ] && extra=""
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2959$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline $extra > $out/bench_n$N.json 2> $out/bench_n$N.err
  show $out/bench_n$N.json n$N | tee -a $out/summary.txt
  tail -2 $out/bench_n$N.err
done
echo done

#!/bin/bash
# 2-GPU call: halo exchange of the gather pass (parity vs the oracle, 1-vs-2 GPU agreement, C5 bench halo on/off), partitioned Lanczos
set -u
out=gpurun_out/r2_multi1
mkdir -p $out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export -f run2
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $out/pytest_multi.log 2>&1
echo "pytest multi rc=$?" | tee $out/rc.txt
tail -25 $out/pytest_multi.log
timeout 400 python scripts/check_multigpu.py 2 > $out/check_1v2.log 2>&1
echo "check 1v2 rc=$?" | tee -a $out/rc.txt
tail -12 $out/check_1v2.log
for halo in 1 0; do
  timeout 300 bash -c "run2 2957$halo bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-solve --option halo=$halo" > $out/bench_halo$halo.json 2> $out/bench_halo$halo.err
  python - $out/bench_halo$halo.json halo$halo <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "it/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "comm_ms", round(d["comm_ms_per_step"], 3),
          " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()), "L=%.15g obj=%.15g" % (d["last_iterate"]["L"], d["last_iterate"]["obj"]),
          "lanczos", d.get("lanczos"), "setup", d.get("setup"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done | tee $out/summary.txt
tail -5 $out/bench_halo1.err
timeout 300 bash -c "run2 29561 scripts/check_multigpu_solve.py" > $out/solve_2gpu.json 2> $out/solve_2gpu.err
SDPLRP_LANCZOS_DIST=1 timeout 300 bash -c "run2 29562 scripts/check_multigpu_solve.py" > $out/solve_2gpu_lzdist.json 2> $out/solve_2gpu_lzdist.err
python - <<'PY'
import json
for f in ("solve_2gpu", "solve_2gpu_lzdist"):
    try:
        d = json.loads(open(f"gpurun_out/r2_multi1/{f}.json").read().strip().splitlines()[-1])
        print(f, "dual", d["dual"], "dual_s %.3f" % d["dual_s"], "it400", d["it400"])
    except Exception as e:
        print(f, "FAILED", e)
PY
echo done

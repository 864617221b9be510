#!/bin/bash
# Second GPU call of round 2 (two GPUs): the row-partitioned Lanczos written blind at the end of round 1.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash scripts/r2_multigpu_call.sh'
# The three JSON lines must agree in "dual" (value, lambda_min, steps) to ~1e-9 relative; "dual_s" is the time of the dual
# check (324 Lanczos steps at C5): replicated operator on 1 and 2 GPUs against the partitioned one on 2.
set -u
out=gpurun_out/r2_multigpu
mkdir -p $out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 300 python scripts/check_multigpu_solve.py > $out/solve_1gpu.json 2> $out/solve_1gpu.err
timeout 300 bash -c "$(declare -f run2); run2 29561 scripts/check_multigpu_solve.py" > $out/solve_2gpu.json 2> $out/solve_2gpu.err
SDPLRP_LANCZOS_DIST=1 timeout 300 bash -c "$(declare -f run2); run2 29562 scripts/check_multigpu_solve.py" > $out/solve_2gpu_lzdist.json 2> $out/solve_2gpu_lzdist.err
SDPLRP_LANCZOS_DIST=1 SDPLRP_LANCZOS_BUNDLE=1 timeout 300 bash -c "$(declare -f run2); run2 29564 scripts/check_multigpu_solve.py" > $out/solve_2gpu_lzdist_bundle.json 2> $out/solve_2gpu_lzdist_bundle.err
SDPLRP_LANCZOS_DIST=1 timeout 300 bash -c "$(declare -f run2); run2 29563 scripts/check_multigpu_parity.py" > $out/parity_2gpu_lzdist.log 2>&1
python - <<'PY'
import json
for f in ("solve_1gpu", "solve_2gpu", "solve_2gpu_lzdist", "solve_2gpu_lzdist_bundle"):
    try:
        d = json.loads(open(f"gpurun_out/r2_multigpu/{f}.json").read().strip().splitlines()[-1])
        print(f, "dual", d["dual"], "dual_s %.3f" % d["dual_s"], "it400", d["it400"])
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 $out/parity_2gpu_lzdist.log

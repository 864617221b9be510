#!/bin/bash
# last single-GPU call of the round: default bench line (with the time-to-tolerance solve), smoke, the full-size C5 parity tests,
# launch list of the final iteration
set -u
out=gpurun_out/r2_final2
mkdir -p $out
( time timeout 400 python bench.py --no-cpu-baseline ) > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?" | tee $out/rc.txt
python - $out/bench_default.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("default it/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}({b['frac']:.2f})" for a, b in k.items()))
print("roofline", {kk: vv for kk, vv in d["roofline"].items() if kk != "kernels"})
print("lanczos", d["lanczos"], "ttt", d["time_to_tol"]); print("clocks", d["clocks"], "launches", d["gpu_launches"], "setup", d["setup"])
PY
tail -3 $out/bench_default.err
( time python -c 'import __graft_entry__ as g; g.smoke()' ) > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/rc.txt; tail -4 $out/smoke.log
( time timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "c5_against or drift" ) > $out/pytest_c5.log 2>&1; echo "pytest c5 rc=$?" | tee -a $out/rc.txt; tail -4 $out/pytest_c5.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-solve --lanczos 5"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_rows|k_gram|k_A_rowc|k_step_grad|k_tail_rest|k_biquadratic|k_obj_slots" -c 96 \
    --csv --log-file $out/launches.csv $B > $out/ncu_launches.log 2>&1
echo "launch list rc=$?" | tee -a $out/rc.txt
ls -la $out
